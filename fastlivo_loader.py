"""Import helper: the package directory is ``fast-livo_b200`` (hyphen), which Python
cannot import by name.  ``load()`` registers it as module ``fastlivo_b200``."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    name = "fastlivo_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg = os.path.join(_ROOT, "fast-livo_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def oracle():
    """TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline): the CPU checker."""
    p = os.path.join(_ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import pyoracle
    return pyoracle

"""CPU tier: the C-ABI library loads without a GPU and exports every entry point include/fastlivo_b200.h declares;
the ctypes binding's symbol list is the header's."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fastlivo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(flb):
    names = _declared()
    assert len(names) >= 45
    L = ctypes.CDLL(flb.build())          # loads on a box without a GPU (no compute call is made)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.flb_abi_version() >= 1


def test_binding_list_matches_header(flb):
    assert sorted(flb.capi.SYMBOLS) == _declared()


def test_create_without_gpu_fails_loudly(flb):
    from conftest import has_gpu
    if has_gpu():
        return
    try:
        flb.Handle(device=0)
    except flb.FlbError as e:
        assert e.code == -3          # FLB_ERR_NO_DEVICE: no CPU fallback
    else:
        raise AssertionError("Handle() must fail without a CUDA device")

"""bench.py runs its CPU-path legs in child processes (the reference's KD_TREE has crashed a process holding several trees):
the children's JSON protocol, and that a child's frame equals the same frame computed in this process."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fastlivo_loader  # noqa: E402


def test_cpu_one_frame_child_matches_in_process():
    r = bench.cpu_one_frame("T1")
    assert "error" not in r, r
    assert r["dt"] > 0 and r["cores"] >= 1 and len(r["x"]) == 24
    flb = fastlivo_loader.load()
    po = fastlivo_loader.oracle()
    frame = flb.synth.make_frame(flb.synth.CONFIGS["T1"])
    run, kind = bench.cpu_frame_runner(po, frame, r["cores"])
    x, rows = run()
    assert kind == r["kind"] and rows > 0
    assert np.array_equal(np.array(r["x"]), x.vector())


def test_cpu_baseline_child_protocol():
    r = bench.cpu_baseline_child("T1", 1)
    assert "error" not in r, r
    assert r["dt4"] > 0 and r["dtall"] > 0 and len(r["x"]) == 24


def test_child_failure_is_reported_not_raised():
    r = bench.cpu_one_frame("no-such-workload")
    assert "error" in r

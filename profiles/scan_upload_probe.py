"""Profiling aid: device time of flb_scan_upload (copy + ordering) for the two ordering paths (FLB_BLOCK_SORT=0/1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fastlivo_loader
flb = fastlivo_loader.load()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
h = flb.Handle(device=0, cell_size=0.6)
h.set_stream(stream.cuda_stream)
for n in (8000, 16000, 24000, 40000):
    f = flb.synth.make_frame("C2")
    scan = h.pinned_like(np.ascontiguousarray(f["scan_body"][:n] if n <= len(f["scan_body"]) else np.tile(f["scan_body"], (2, 1))[:n], np.float32))
    ts = []
    for it in range(25):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record(stream)
        h.scan_upload(scan)
        b.record(stream)
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b) * 1e3)
    print("n", n, "block_sort", os.environ.get("FLB_BLOCK_SORT", "1"), "us: median %.1f min %.1f" % (float(np.median(ts[5:])), float(np.min(ts[5:]))))
h.close()

"""Debugging aid: the secondary legs of bench.py one by one, with faulthandler."""
import faulthandler, json, os, sys
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import fastlivo_loader
flb = fastlivo_loader.load()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
hbm, src = bench.peaks()
po = fastlivo_loader.oracle() if "--oracle" in sys.argv else None
for name in ("C3", "C4"):
    print("other", name, flush=True)
    r = bench.measure_other_workload(flb, torch, name, 0, dev, stream, flush, hbm, src, po)
    print(name, round(r["value"], 1), flush=True)
for name, B in (("C2", 64), ("C4", 16)):
    print("batched", name, B, flush=True)
    r = bench.measure_batched(flb, torch, name, B, 0, dev, stream, flush, hbm, src)
    print(name, B, round(r["value"], 1), flush=True)
print("full frame", flush=True)
r = bench.measure_full_frame(flb, torch, "C2", 0, dev, stream, flush)
print(round(r["value"], 1), flush=True)

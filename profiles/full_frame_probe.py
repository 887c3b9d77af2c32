"""Profiling aid: the e2e_full_frame leg of bench.py alone (env switches of the library apply: FLB_DEFER, FLB_RESERVE_SM, ...)."""
import json, os, sys
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
raw = []
r = bench.measure_full_frame(flb, torch, "C2", 0, dev, stream, flush, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 20, raw=raw)
print(json.dumps({k: r[k] for k in ("value", "ms_per_frame", "host_ms_per_stage")}))
for k, row in enumerate(raw):
    print(k, row)

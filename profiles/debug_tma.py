import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, fastlivo_loader
flb = fastlivo_loader.load()
f = flb.synth.make_frame("T0")
h = flb.Handle(device=0)
h.load_frame(f)
x = flb.capi.State18.from_frame(f)
try:
    rep = h.vio_update(flb.capi.vio_params(f, 3), x, x.copy())
    print("TMA run ok", list(rep.passes), x.vector()[:12])
except Exception as e:
    print("TMA run failed:", e)

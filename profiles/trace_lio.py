"""Profiling aid: device-side stage timestamps of the persistent LIO kernel (C2 workload)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import fastlivo_loader
flb = fastlivo_loader.load()
f = flb.synth.make_frame("C2")
h = flb.Handle(cell_size=0.6)
h.load_frame(f)
lprm = flb.capi.lio_params(f, 2, early_stop=False)
x = flb.capi.State18.from_frame(f)
for it in range(3):
    h.trace_enable(True)
    h.state_upload(x, x.copy()); h.lio_update_enqueue(lprm); h.synchronize()
    us = np.zeros(127); n = C.c_int()
    h._ck(h.L.flb_trace_download(h.h, 0, us.ctypes.data_as(C.c_void_p), 127, C.byref(n)))
t = np.concatenate([[0], us])
print("pass/solve", np.round(np.diff(t[:7]), 2))
for p in range(3):
    fine = t[64 + 8 * p:64 + 8 * p + 5]
    print("pass", p, "arrive", round(t[1 + 2 * p], 2), "leader stamps rel arrive:", np.round(fine - t[1 + 2 * p], 2))
pr = t[112:116]
print("probe thread (last rematch pass): start", round(pr[0], 2), "knn", round(pr[1] - pr[0], 2), "plane", round(pr[2] - pr[1], 2),
      "residual+row", round(pr[3] - pr[2], 2), "; pass began at", round(t[4], 2), "all arrived at", round(t[5], 2))

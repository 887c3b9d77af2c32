"""Profiling aid: device-side stage timestamps of the persistent VIO kernel (C2 workload)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import fastlivo_loader
flb = fastlivo_loader.load()
f = flb.synth.make_frame("C2")
h = flb.Handle(cell_size=0.6)
h.load_frame(f)
vprm = flb.capi.vio_params(f, 3, force_all_passes=True)
x = flb.capi.State18.from_frame(f)
for it in range(3):
    h.trace_enable(True)
    h.state_upload(x, x.copy()); h.vio_update_enqueue(vprm); h.synchronize()
    us = np.zeros(127); n = C.c_int()
    h._ck(h.L.flb_trace_download(h.h, 1, us.ctypes.data_as(C.c_void_p), 127, C.byref(n)))
t = np.concatenate([[0], us])
print("pass/solve", np.round(np.diff(t[:19]), 2))
names = ["errs staged", "reduce done", "step done", "err-sum done", "joined", "ctrl done", "state done"]
for p in range(9):
    fine = t[32 + 8 * p:32 + 8 * p + 7]
    print("pass", p, "arrive %.2f" % t[1 + 2 * p], "published +%.2f" % (t[2 + 2 * p] - t[1 + 2 * p]), "leader stamps rel arrive:",
          " ".join("%s %.2f" % (nm, v - t[1 + 2 * p]) for nm, v in zip(names, fine)))
h.L.flb_debug_vio_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
buf = np.zeros((512 * 64,), np.uint64); nb = C.c_int(); wpb = C.c_int()
h.L.flb_debug_vio_stamps(h.h, buf.ctypes.data_as(C.c_void_p), 512, C.byref(nb), C.byref(wpb))
b = buf[:nb.value * wpb.value].reshape(nb.value, wpb.value).astype(np.float64)
t0 = b[:, 0][b[:, 0] > 0].min()
rel = np.where(b > 0, (b - t0) * 1e-3, np.nan)
def st(x):
    x = x[~np.isnan(x)]
    return "n %d min %.2f p50 %.2f p90 %.2f max %.2f" % (len(x), x.min(), np.percentile(x, 50), np.percentile(x, 90), x.max())
print("blocks", nb.value, "(stamps of the LAST pass)")
print("wake (loop start)   ", st(rel[:, 0]))
print("pose ready          ", st(rel[:, 1]))
w = rel[:, 8:].reshape(nb.value, -1, 4)
print("warp geom done      ", st(w[:, :, 0]))
print("warp taps staged    ", st(w[:, :, 1]))
print("warp pixels done    ", st(w[:, :, 2]))
print("warp err chain done ", st(w[:, :, 3]))
print("block reduce stored ", st(rel[:, 2]))
print("arrived             ", st(rel[:, 3]))

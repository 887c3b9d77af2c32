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
# per-block stamps of the last pass
h.L.flb_debug_block_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
buf = np.zeros((512, 16), np.uint64); nb = C.c_int()
h.L.flb_debug_block_stamps(h.h, buf.ctypes.data_as(C.c_void_p), 512, C.byref(nb))
b = buf[:nb.value - 1].astype(np.float64)   # worker blocks
t0 = b[:, 0].min()
rel = (b - t0) * 1e-3
def st(x): return "min %.2f p50 %.2f p90 %.2f max %.2f" % (x.min(), np.percentile(x, 50), np.percentile(x, 90), x.max())
print("blocks", nb.value)
print("wake (loop start)   ", st(rel[:, 0]))
print("pose ready          ", st(rel[:, 1]))
w = rel[:, 4:12]; w = w[w > 0]
print("warp compute done   ", st(w))
print("block reduce stored ", st(rel[:, 2]))
print("arrived             ", st(rel[:, 3]))
ww = rel[:, 4:12]
for wi in range(8):
    col = ww[:, wi]; col = col[col > 0]
    if len(col): print("warp slot", wi, "n", len(col), st(col))
nwork = nb.value - 1
flat = [(ww[bi, wi], bi, wi, wi * nwork + bi) for bi in range(nwork) for wi in range(8) if ww[bi, wi] > 0]
flat.sort(reverse=True)
print("slowest warps (done_us, block, warp, chunk):", [(round(a, 1), b_, w_, c_) for a, b_, w_, c_ in flat[:16]])
chunks = np.array([c_ for _, _, _, c_ in flat]); times = np.array([a for a, _, _, _ in flat])
order = np.argsort(chunks)
# smooth time vs chunk id (Morton order => spatial position)
ts = times[order]
print("mean done time by chunk decile:", [round(float(x.mean()), 1) for x in np.array_split(ts, 10)])

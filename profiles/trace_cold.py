"""Profiling aid: per-block stage stamps of the FIRST LIO pass with a cold (flushed) L2, C2 workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import fastlivo_loader
flb = fastlivo_loader.load()
f = flb.synth.make_frame("C2")
h = flb.Handle(cell_size=0.6)
h.load_frame(f)
lprm = flb.capi.lio_params(f, 0, early_stop=False)     # T = 0: exactly one pass (iterCount = -1)
x = flb.capi.State18.from_frame(f)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
h.L.flb_debug_block_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
for cold in (False, True, True):
    for it in range(2):
        h.trace_enable(True)
        h.state_upload(x, x.copy()); h.synchronize()
        if cold:
            flush.zero_(); torch.cuda.synchronize()
        h.lio_update_enqueue(lprm); h.synchronize()
    us = np.zeros(127); n = C.c_int()
    h._ck(h.L.flb_trace_download(h.h, 0, us.ctypes.data_as(C.c_void_p), 127, C.byref(n)))
    t = np.concatenate([[0], us])
    fine = t[64:69]
    buf = np.zeros((512, 16), np.uint64); nb = C.c_int()
    h.L.flb_debug_block_stamps(h.h, buf.ctypes.data_as(C.c_void_p), 512, C.byref(nb))
    b = buf[:nb.value - 1].astype(np.float64)           # worker blocks
    # absolute ns -> us relative to the kernel-start stamp: trace[0] is absolute too, but downloaded relative;
    # use the earliest worker wake as origin
    t0 = b[:, 0].min()
    rel = (b - t0) * 1e-3
    def st(v): return "min %.2f p50 %.2f p90 %.2f max %.2f" % (v.min(), np.percentile(v, 50), np.percentile(v, 90), v.max())
    print("=== cold L2" if cold else "=== warm L2")
    print("all arrived (leader) %.2f  leader stamps rel arrive %s" % (t[1], np.round(fine - t[1], 2)))
    print("wake        ", st(rel[:, 0]))
    print("pose ready  ", st(rel[:, 1]))
    w = rel[:, 4:12]; print("warp done   ", st(w[w > 0]))
    print("stored      ", st(rel[:, 2]))
    print("arrived     ", st(rel[:, 3]))

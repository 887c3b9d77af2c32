import sys; sys.path.insert(0,'/root/repo')
import numpy as np, fastlivo_loader, ctypes as C
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
print("pass/solve", np.round(np.diff(t[:7]),2))
for p in range(3):
    fine = t[64+8*p:64+8*p+5]
    print("pass", p, "arrive", round(t[1+2*p],2), "fine stamps rel arrive:", np.round(fine - t[1+2*p], 2), "release", round(t[2+2*p]-t[1+2*p],2))

"""Profiling aid: host wall time of each C-ABI call of one end-to-end frame (C2 workload, host buffers)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fastlivo_loader
flb = fastlivo_loader.load()
f = flb.synth.make_frame("C2")
h = flb.Handle(cell_size=0.6)
h.load_frame(f)
lprm = flb.capi.lio_params(f, 2, early_stop=False)
vprm = flb.capi.vio_params(f, 3, force_all_passes=True)
x0 = flb.capi.State18.from_frame(f)
scan, img = f["scan_body"], f["image"]
ppos, pref, plev = f["patch_pos"], f["patch_ref"], f["patch_level"]
names = ["scan_upload", "image_upload", "patches_upload", "lio_update", "vio_update"]
T = {n: [] for n in names}
for it in range(120):
    t = [time.perf_counter()]
    h.scan_upload(scan); t.append(time.perf_counter())
    h.image_upload(img); t.append(time.perf_counter())
    h.patches_upload(ppos, pref, plev); t.append(time.perf_counter())
    x = x0.copy()
    h.lio_update(lprm, x, x0); t.append(time.perf_counter())
    xp = x.copy()
    h.vio_update(vprm, x, xp); t.append(time.perf_counter())
    if it >= 20:
        for n, a, b in zip(names, t[:-1], t[1:]):
            T[n].append(b - a)
tot = 0.0
for n in names:
    m = 1e6 * float(np.median(T[n])); tot += m
    print("%-16s %8.1f us" % (n, m))
print("%-16s %8.1f us  -> %.0f frames/s" % ("sum", tot, 1e6 / tot))

"""Generate tests/golden/T0_golden.npz: a small self-contained frame (inputs) plus the oracle's
outputs on it.  Run from the repo root:  python tests/golden/make_golden.py

"parity unpinned": the reference has no golden vectors of its own and cannot be
built here, so these come from OUR oracle (oracle/flo_oracle.cpp; kNN cross-checked against the
reference's ikd_Tree.cpp in tests/test_oracle_properties.py).  They pin the oracle against
compiler / platform drift and give the GPU tier a fixture that does not depend on the generator.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fastlivo_loader  # noqa: E402


def main():
    flb = fastlivo_loader.load()
    po = fastlivo_loader.oracle()
    f = flb.synth.make_frame("T0")
    out = {k: f[k] for k in ("map_xyz", "scan_body", "R_prop", "p_prop", "cov", "vel", "bg", "ba", "grav", "R_LI", "t_LI",
                             "Rcl", "Pcl", "image", "patch_pos", "patch_ref", "patch_level")}
    cam = f["cam"]
    out["cam"] = np.array([cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], *cam["d"]])
    out["cfg"] = np.array([f["cfg"].laser_point_cov, f["cfg"].img_point_cov, f["cfg"].cell_size])
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    o = lio.run_pass(po.lio_params(f, 3), f["R_prop"], f["p_prop"], True, rows12=True)
    for k in ("world", "nn_idx", "nn_d2", "pabcd", "pd2", "selected", "Hsub", "h_x", "meas", "sel_idx", "HTH6", "HTz6",
              "HTH12", "HTh12"):
        out["lio_" + k] = o[k]
    x = po.state_from_frame(f)
    rep = lio.update(po.lio_params(f, 4), x, x.copy())
    out["lio_state"] = x.vector()
    out["lio_cov"] = x.P
    out["lio_report"] = np.array([rep.passes, rep.knn_passes, rep.n_eff_last, rep.rows_total])
    vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
    for level in (2, 0):
        v = vio.run_pass(po.vio_params(f, 3), f["R_prop"], f["p_prop"], level)
        out[f"vio{level}_z"] = v["z"]
        out[f"vio{level}_H"] = v["H_sub"][:8 * 64]          # first 8 patches (keeps the fixture small)
        out[f"vio{level}_errors"] = v["errors"]
        out[f"vio{level}_HTH6"] = v["HTH6"]
        out[f"vio{level}_HTz6"] = v["HTz6"]
        out[f"vio{level}_scalar"] = np.array([v["error"], v["n_meas"], v["skipped"]], np.float64)
    xv = x.copy()
    vrep = vio.update(po.vio_params(f, 4), xv, x.copy())
    out["vio_state"] = xv.vector()
    out["vio_cov"] = xv.P
    out["vio_report"] = np.array([*vrep.passes, vrep.rows_total, vrep.cov_updated])
    out["vio_last_error"] = np.array(list(vrep.last_error), np.float32)
    path = os.path.join(ROOT, "tests", "golden", "T0_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Generate tests/golden/imu_golden.npz: small ImuProcess::UndistortPcl cases (inputs from the seeded generator)
plus the oracle's outputs on them.  Run from the repo root:  python tests/golden/make_golden_imu.py

"parity unpinned": the reference has no fixtures for this step; these pin OUR oracle
(oracle/flo_imu.cpp, cross-checked against a literal Python replay and closed forms in tests/test_imu_oracle.py)
against compiler / platform drift and give the GPU tier a fixture that does not depend on the generator.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fastlivo_loader  # noqa: E402
from imu_util import oracle_inputs  # noqa: E402

VARIANTS = ["nominal", "late_imu", "imu_past_end", "stale_imu", "imu_before_scan", "unsorted_points", "early_points"]


def main():
    flb = fastlivo_loader.load()
    po = fastlivo_loader.oracle()
    out = {}
    for v in VARIANTS:
        f = flb.synth.make_imu_frame(seed=11, n_points=600, variant=v)
        P, C, x = oracle_inputs(po, f)
        pts, poses = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
        for k in ("v_imu", "pts", "offset_ms"):
            out[f"{v}_in_{k}"] = f[k]
        out[f"{v}_in_times"] = np.array([f["pcl_beg_time"], f["pcl_end_time"], f["last_lidar_end_time"]])
        out[f"{v}_pts"] = pts
        out[f"{v}_poses"] = poses
        out[f"{v}_state"] = np.concatenate([x.rot[:], x.pos[:], x.vel[:]])
        out[f"{v}_cov"] = np.array(x.cov[:]).reshape(18, 18)
        out[f"{v}_carry"] = np.concatenate([[C.last_lidar_end_time], C.acc_s_last[:], C.angvel_last[:]])
    path = os.path.join(ROOT, "tests", "golden", "imu_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

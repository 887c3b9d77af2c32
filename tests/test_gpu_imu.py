"""GPU tier, SURVEY.md section 8 row f3: flb_imu_undistort (IMU forward propagation of state + 18x18 covariance,
frame-end prediction, per-point backward undistortion) against the oracle restatement of
ImuProcess::UndistortPcl (reference src/IMU_Processing.cpp:655-808), through the C ABI.

Tolerances: double-precision state / covariance / IMUpose to 1e-12 relative (sin / cos of libm vs CUDA differ in
the last bit); compensated points are float32 roundings of double results, so they must be bit-identical except
where that last bit flips a rounding (allowed: < 0.1 % of coordinates, each within 1e-5 m)."""
import os

import numpy as np
import pytest

from imu_util import oracle_inputs, product_inputs

pytestmark = pytest.mark.gpu

VARIANTS = ["nominal", "late_imu", "imu_past_end", "stale_imu", "imu_before_scan", "unsorted_points", "early_points"]


def _check(flb, po, f):
    P, C, x = oracle_inputs(po, f)
    ref_pts, ref_poses = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    Pg, Cg, xg = product_inputs(flb, f)
    h = flb.Handle(device=0)
    h.state_upload(xg, xg.copy())
    packed = np.concatenate([f["pts"], f["offset_ms"][:, None]], 1).astype(np.float32)
    got_pts, got_poses = h.imu_undistort(Pg, Cg, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], packed, offset_index=3)
    xs, _, _ = h.state_download()
    assert got_poses.shape == ref_poses.shape
    assert np.allclose(got_poses, ref_poses, rtol=1e-12, atol=1e-12)
    for name in ("rot", "pos", "vel", "bg", "ba", "grav"):
        assert np.allclose(np.array(getattr(xs, name)[:]), np.array(getattr(x, name)[:]), rtol=1e-12, atol=1e-12), name
    cg, cr = np.array(xs.cov[:]).reshape(18, 18), np.array(x.cov[:]).reshape(18, 18)
    assert np.abs(cg - cr).max() <= 1e-12 * np.abs(cr).max()
    assert Cg.last_lidar_end_time == C.last_lidar_end_time
    assert np.allclose(Cg.acc_s_last[:], C.acc_s_last[:], rtol=1e-12, atol=1e-12)
    assert np.allclose(Cg.angvel_last[:], C.angvel_last[:], rtol=1e-12, atol=1e-12)
    if len(ref_pts):
        assert np.abs(got_pts - ref_pts).max() <= 1e-5
        assert (got_pts == ref_pts).mean() > 0.999
    return h


@pytest.mark.parametrize("variant", VARIANTS)
def test_imu_undistort_matches_oracle(flb, po, variant):
    _check(flb, po, flb.synth.make_imu_frame(seed=21, n_points=24000, variant=variant))


def test_imu_undistort_full_size_and_strided_points(flb, po):
    """100 k points in the reference's own point layout (pcl::PointXYZINormal: stride 12, curvature at 9)."""
    f = flb.synth.make_imu_frame(seed=22, n_points=100000)
    P, C, x = oracle_inputs(po, f)
    ref_pts, _ = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    Pg, Cg, xg = product_inputs(flb, f)
    h = flb.Handle(device=0)
    h.state_upload(xg, xg.copy())
    pcl = np.zeros((len(f["pts"]), 12), np.float32)
    pcl[:, :3] = f["pts"]; pcl[:, 3] = 1.0; pcl[:, 8] = 17.0; pcl[:, 9] = f["offset_ms"]
    got, _ = h.imu_undistort(Pg, Cg, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], pcl, offset_index=9)
    assert np.abs(got - ref_pts).max() <= 1e-5 and (got == ref_pts).mean() > 0.999


def test_imu_undistort_degenerate_inputs(flb, po):
    f = flb.synth.make_imu_frame(seed=23, n_points=0)
    _check(flb, po, f)                                   # no points: propagation only (:776)
    f = flb.synth.make_imu_frame(seed=24, n_points=1)
    _check(flb, po, f)
    f = flb.synth.make_imu_frame(seed=25, n_points=50)
    f["v_imu"] = f["v_imu"][:1]                          # only last_imu_: no interval, prediction from the carried rates
    _check(flb, po, f)


def test_imu_undistort_golden(flb):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "imu_golden.npz"))
    for variant in VARIANTS:
        f = flb.synth.make_imu_frame(seed=11, n_points=600, variant=variant)
        Pg, Cg, xg = product_inputs(flb, f)
        h = flb.Handle(device=0)
        h.state_upload(xg, xg.copy())
        packed = np.concatenate([g[f"{variant}_in_pts"], g[f"{variant}_in_offset_ms"][:, None]], 1).astype(np.float32)
        t = g[f"{variant}_in_times"]
        Cg.last_lidar_end_time = t[2]
        got, poses = h.imu_undistort(Pg, Cg, g[f"{variant}_in_v_imu"], t[0], t[1], packed, offset_index=3)
        assert np.abs(got - g[f"{variant}_pts"]).max() <= 1e-5 and (got == g[f"{variant}_pts"]).mean() > 0.999
        assert np.allclose(poses, g[f"{variant}_poses"], rtol=1e-12, atol=1e-12)


def test_imu_undistort_rejects_bad_input(flb):
    f = flb.synth.make_imu_frame(seed=26, n_points=10)
    Pg, Cg, xg = product_inputs(flb, f)
    h = flb.Handle(device=0)
    packed = np.concatenate([f["pts"], f["offset_ms"][:, None]], 1).astype(np.float32)
    with pytest.raises(flb.capi.FlbError):               # no device state yet
        h.imu_undistort(Pg, Cg, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], packed)
    h.state_upload(xg, xg.copy())
    bad = f["v_imu"].copy(); bad[2, 0] = bad[1, 0] - 1.0
    with pytest.raises(flb.capi.FlbError):               # IMU time running backwards
        h.imu_undistort(Pg, Cg, bad, f["pcl_beg_time"], f["pcl_end_time"], packed)

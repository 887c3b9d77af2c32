"""CPU tier, row f3: properties of the oracle restatement of ImuProcess::UndistortPcl
(reference src/IMU_Processing.cpp:655-808) -- the reference ships no fixtures for this step."""
import numpy as np
import pytest

from imu_util import oracle_inputs, exp_w_dt

VARIANTS = ["nominal", "late_imu", "imu_past_end", "stale_imu", "imu_before_scan", "unsorted_points", "early_points"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_runs_and_keeps_covariance_symmetric_psd(flb, po, variant):
    f = flb.synth.make_imu_frame(seed=3, n_points=4000, variant=variant)
    P, C, x = oracle_inputs(po, f)
    pts, poses = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    cov = np.array(x.cov[:]).reshape(18, 18)
    assert np.isfinite(cov).all() and np.isfinite(pts).all()
    assert np.abs(cov - cov.T).max() <= 1e-12 * np.abs(cov).max()
    assert np.linalg.eigvalsh(0.5 * (cov + cov.T)).min() > -1e-12
    R = np.array(x.rot[:]).reshape(3, 3)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12
    assert C.last_lidar_end_time == f["pcl_end_time"]                       # :761
    assert poses[0, 0] == 0.0 and len(poses) >= 1


def test_no_motion_no_compensation(flb, po):
    """Zero rates, gravity-cancelling acceleration, zero velocity: every pose equals the start pose, so the
    compensation is the identity up to float rounding of the double round trip."""
    f = flb.synth.make_imu_frame(seed=5, n_points=3000)
    f["v_imu"][:, 1:4] = 0.0
    f["v_imu"][:, 4:7] = np.array([0.0, 0.0, 9.81]) * f["mean_acc_norm"] / f["G_m_s2"]
    f["bg"][:] = 0; f["ba"][:] = 0; f["vel"][:] = 0; f["R"] = np.eye(3)
    f["acc_s_last"][:] = 0; f["angvel_last"][:] = 0
    P, C, x = oracle_inputs(po, f)
    pts, _ = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    assert np.abs(pts - f["pts"]).max() < 2e-5
    assert np.abs(np.array(x.pos[:]) - f["p"]).max() < 1e-9 and np.abs(np.array(x.vel[:])).max() < 1e-9


def test_constant_rate_compensation_matches_closed_form(flb, po):
    """Constant body rate, no translation: a point at time t must be rotated by Exp(w (t - t_end)) (in the IMU
    frame), whatever the IMU sampling -- checks the pose bookkeeping and the backward pass end to end."""
    f = flb.synth.make_imu_frame(seed=7, n_points=2000)
    w = np.array([0.3, -0.2, 0.5])
    f["v_imu"][:, 1:4] = w
    f["v_imu"][:, 4:7] = np.array([0.0, 0.0, 9.81]) * f["mean_acc_norm"] / f["G_m_s2"]
    f["bg"][:] = 0; f["ba"][:] = 0; f["vel"][:] = 0; f["R"] = np.eye(3); f["p"][:] = 0
    f["grav"] = np.array([0.0, 0.0, -9.81])
    f["acc_s_last"][:] = 0; f["angvel_last"] = w.copy()
    f["R_LI"] = np.eye(3); f["t_LI"] = np.zeros(3)
    # gravity is only cancelled while the attitude is the identity; make translation irrelevant instead:
    f["v_imu"][:, 4:7] = 0.0; f["grav"] = np.zeros(3)
    P, C, x = oracle_inputs(po, f)
    pts, _ = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    t_end = f["pcl_end_time"] - f["pcl_beg_time"]
    t = f["offset_ms"].astype(np.float64) / 1000.0
    exp = np.stack([exp_w_dt(w, ti - t_end) @ pi for ti, pi in zip(t, f["pts"].astype(np.float64))])
    sel = t > 0
    assert np.abs(pts[sel] - exp[sel]).max() < 5e-4          # piecewise re-anchoring at the IMU poses + float storage


def test_first_point_is_compensated_once_per_earlier_pose(flb, po):
    """The reference's loop exit (:807) leaves the first point to be re-tested against every earlier IMU pose.
    With a late first point the oracle must reproduce that: compare with a literal Python replay."""
    f = flb.synth.make_imu_frame(seed=9, n_points=300, variant="unsorted_points")
    P, C, x = oracle_inputs(po, f)
    pts, poses = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    R_LI, t_LI = np.array(f["R_LI"]), np.array(f["t_LI"])
    R_end, p_end = np.array(x.rot[:]).reshape(3, 3), np.array(x.pos[:])
    ext, ext_t = R_LI.T @ R_end.T, R_LI.T @ t_LI
    q = f["pts"].astype(np.float32).copy()
    t = f["offset_ms"].astype(np.float64) / 1000.0
    ip = len(q) - 1
    for kp in range(len(poses) - 1, 0, -1):
        h = poses[kp - 1]
        while t[ip] > h[0]:
            dt = t[ip] - h[0]
            R_i = h[13:22].reshape(3, 3) @ exp_w_dt(h[4:7], dt)
            T_ei = h[10:13] + h[7:10] * dt + 0.5 * h[1:4] * dt * dt - p_end
            q[ip] = (ext @ (R_i @ (R_LI @ q[ip].astype(np.float64) + t_LI) + T_ei) - ext_t).astype(np.float32)
            if ip == 0:
                break
            ip -= 1
    assert np.abs(pts - q).max() < 1e-4
    assert np.abs(pts[0] - q[0]).max() < 1e-5


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_matches_golden(flb, po, variant):
    """tests/golden/imu_golden.npz (made by tests/golden/make_golden_imu.py): generator and oracle are stable."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "imu_golden.npz"))
    f = flb.synth.make_imu_frame(seed=11, n_points=600, variant=variant)
    assert np.array_equal(f["v_imu"], g[f"{variant}_in_v_imu"]) and np.array_equal(f["pts"], g[f"{variant}_in_pts"])
    assert np.array_equal(f["offset_ms"], g[f"{variant}_in_offset_ms"])
    P, C, x = oracle_inputs(po, f)
    pts, poses = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    assert np.allclose(pts, g[f"{variant}_pts"], rtol=0, atol=1e-5) and (pts == g[f"{variant}_pts"]).mean() > 0.999
    assert np.allclose(poses, g[f"{variant}_poses"], rtol=1e-13, atol=1e-13)
    assert np.allclose(np.concatenate([x.rot[:], x.pos[:], x.vel[:]]), g[f"{variant}_state"], rtol=1e-13, atol=1e-13)
    assert np.allclose(np.array(x.cov[:]).reshape(18, 18), g[f"{variant}_cov"], rtol=1e-12, atol=1e-18)


def test_device_point_math_matches_oracle_on_the_host(flb, po, hostemu):
    """The product's per-point compensation (flb_device.cuh, compiled for the host) against the oracle: with a single
    IMU interval every point is compensated with pose 0, so the oracle's output isolates that function."""
    import ctypes as C
    f = flb.synth.make_imu_frame(seed=13, n_points=5000)
    f["v_imu"] = f["v_imu"][:1]                       # no interval: IMUpose = [pose 0], prediction from the carried rates
    P, Cc, x = oracle_inputs(po, f)
    pose0 = np.concatenate([[0.0], f["acc_s_last"], f["angvel_last"], f["vel"], f["p"], np.asarray(f["R"]).ravel()])
    ref, poses = po.imu_undistort(P, Cc, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    assert len(poses) == 1
    # with one pose there is no head/tail pair, so the reference's backward loop does not run at all:
    assert np.array_equal(ref, f["pts"])
    # drive the device function directly with pose 0 and compare against the closed form in double
    out = np.zeros_like(f["pts"])
    _p = lambda a: a.ctypes.data_as(C.c_void_p)
    R_LI = np.ascontiguousarray(f["R_LI"], np.float64); t_LI = np.ascontiguousarray(f["t_LI"], np.float64)
    rot_end = np.array(x.rot[:]); pos_end = np.array(x.pos[:])
    hostemu.emu_imu_compensate(_p(np.ascontiguousarray(pose0)), _p(R_LI), _p(t_LI), _p(rot_end), _p(pos_end),
                               _p(np.ascontiguousarray(f["pts"])), _p(np.ascontiguousarray(f["offset_ms"])), len(out), _p(out))
    t = f["offset_ms"].astype(np.float64) / 1000.0
    Rend = rot_end.reshape(3, 3)
    ext, ext_t = R_LI.T @ Rend.T, R_LI.T @ t_LI
    exp = np.stack([ext @ (np.asarray(f["R"]) @ exp_w_dt(f["angvel_last"], ti) @ (R_LI @ q + t_LI)
                           + (f["p"] + f["vel"] * ti + 0.5 * f["acc_s_last"] * ti * ti - pos_end)) - ext_t
                    for ti, q in zip(t, f["pts"].astype(np.float64))])
    assert np.abs(out - exp).max() < 2e-5
    # and Exp(w, dt) itself, against the oracle's poses of a propagated frame (rot = R0 * Exp(...) chain is covered there)
    E = np.zeros(9)
    hostemu.emu_exp_w_dt(_p(np.array([0.3, -0.2, 0.5])), C.c_double(0.01), _p(E))
    assert np.abs(E.reshape(3, 3) - exp_w_dt(np.array([0.3, -0.2, 0.5]), 0.01)).max() < 1e-15

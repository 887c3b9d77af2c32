"""Helpers shared by the IMU propagation / undistortion tests."""
import numpy as np


def _fill(P, C, f):
    P.cov_gyr[:] = f["cov_gyr"]; P.cov_acc[:] = f["cov_acc"]
    P.cov_bias_gyr[:] = f["cov_bias_gyr"]; P.cov_bias_acc[:] = f["cov_bias_acc"]
    P.G_m_s2 = f["G_m_s2"]; P.mean_acc_norm = f["mean_acc_norm"]
    P.R_LI[:] = np.asarray(f["R_LI"], np.float64).ravel(); P.t_LI[:] = f["t_LI"]
    C.last_lidar_end_time = f["last_lidar_end_time"]
    C.acc_s_last[:] = f["acc_s_last"]; C.angvel_last[:] = f["angvel_last"]
    return P, C


def oracle_inputs(po, f):
    P, C = _fill(po.ImuParams(), po.ImuCarry(), f)
    x = po.State18.make(f["R"], f["p"], f["vel"], f["bg"], f["ba"], f["grav"], f["cov"])
    return P, C, x


def product_inputs(flb, f):
    P, C = _fill(flb.capi.ImuParams(), flb.capi.ImuCarry(), f)
    x = flb.capi.State18.make(f["R"], f["p"], f["vel"], f["bg"], f["ba"], f["grav"], f["cov"])
    return P, C, x


def exp_w_dt(w, dt):
    n = np.linalg.norm(w)
    if not n > 1e-7:
        return np.eye(3)
    a = w / n
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    ang = n * dt
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K

"""CPU tier: the algebra behind the leader's per-pass solve ("The solve").

Reference form (src/laserMapping.cpp:1664-1683, src/lidar_selection.cpp:871-879), H^T H non-zero only in its leading
6x6 block:
    K1 = (HTH18 + (P / sigma)^-1)^-1 ;  G = K1 HTH18 ;  solution = sign K1[:, :6] HTz + vec - G vec
Product form (leader_fast_solve / leader_gain, flb_kernels.cuh):
    Kt = (HTH6 + sigma P11^-1)^-1 ; B = P21 P11^-1 ; y = Kt (sign HTz - HTH6 vec6) ; solution = vec + [y; B y]
    and, for the covariance update only, G[:, :6] = [Kt; B Kt] HTH6.
Both must agree to rounding for any SPD covariance."""
import numpy as np
import pytest


def _spd(rng, n, scale):
    a = rng.normal(size=(n, n))
    return (a @ a.T + n * np.eye(n)) * scale


@pytest.mark.parametrize("sign", [+1.0, -1.0])
@pytest.mark.parametrize("seed", range(6))
def test_single_rhs_form_equals_reference_form(sign, seed):
    rng = np.random.default_rng(seed)
    P = _spd(rng, 18, 10.0 ** rng.uniform(-6, -2))
    H = rng.normal(size=(200, 6)) * 10.0 ** rng.uniform(-1, 2)
    z = rng.normal(size=200)
    sigma = 10.0 ** rng.uniform(-3, 2)
    vec = rng.normal(size=18) * 1e-2
    HTH6, HTz6 = H.T @ H, H.T @ z
    HTH18 = np.zeros((18, 18)); HTH18[:6, :6] = HTH6
    HTz18 = np.zeros(18); HTz18[:6] = HTz6
    # reference
    K1 = np.linalg.inv(HTH18 + np.linalg.inv(P / sigma))
    G = K1 @ HTH18
    ref = sign * (K1 @ HTz18) + vec - G @ vec
    # product
    P11inv = np.linalg.inv(P[:6, :6])
    B = P[6:, :6] @ P11inv
    Kt = np.linalg.inv(HTH6 + sigma * P11inv)
    y = Kt @ (sign * HTz6 - HTH6 @ vec[:6])
    got = vec + np.concatenate([y, B @ y])
    assert np.abs(got - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-12)
    Gc = np.vstack([Kt, B @ Kt]) @ HTH6
    assert np.abs(Gc - G[:, :6]).max() <= 1e-9 * np.abs(G).max()
    assert np.abs(G[:, 6:]).max() <= 1e-12 * max(np.abs(G).max(), 1.0)
    # covariance update: (I - G) P == P - Gc P[:6, :]
    assert np.abs((np.eye(18) - G) @ P - (P - Gc @ P[:6, :])).max() <= 1e-9 * np.abs(P).max()

"""GPU tier at BASELINE.json's full sizes (C1..C4): the whole frame against the CPU path (oracle +
the reference's ikd-Tree), plus size-independent properties (sortedness, idempotence, symmetry)."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def _frame_cpu(po, f, tree):
    cfg = f["cfg"]
    lio = po.Lio(f["map_xyz"], f["scan_body"], tree)
    x = po.state_from_frame(f)
    lrep = lio.update(po.lio_params(f, cfg.lio_passes - 1, nthreads=8, early_stop=False), x, x.copy())
    vrep = None
    if cfg.n_patch:
        vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
        vrep = vio.update(po.vio_params(f, cfg.vio_passes, early_stop=False, force_all_passes=True), x, x.copy())
    return x, lrep, vrep


def _frame_gpu(flb, h, f):
    cfg = f["cfg"]
    x = flb.capi.State18.from_frame(f)
    lrep = h.lio_update(flb.capi.lio_params(f, cfg.lio_passes - 1, early_stop=False), x, x.copy())
    vrep = None
    if cfg.n_patch:
        vrep = h.vio_update(flb.capi.vio_params(f, cfg.vio_passes, early_stop=False, force_all_passes=True), x, x.copy())
    return x, lrep, vrep


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
def test_full_size_frame_matches_cpu_path(flb, po, name):
    if po.ref_lib() is None:
        pytest.skip("needs oracle/_ref/libikdtree_ref.so (brute-force kNN is too slow at this size)")
    f = flb.synth.make_frame(name)
    cfg = f["cfg"]
    tree = po.IkdTreeRef(f["map_xyz"])
    xo, lo, vo = _frame_cpu(po, f, tree)
    h = flb.Handle(device=0, cell_size=cfg.cell_size)
    h.load_frame(f)
    xg, lg, vg = _frame_gpu(flb, h, f)
    assert (lg.passes, lg.knn_passes, lg.n_eff_last, lg.rows_total) == (lo.passes, lo.knn_passes, lo.n_eff_last, lo.rows_total)
    if vo is not None:
        assert list(vg.passes) == list(vo.passes) and vg.rows_total == vo.rows_total
        np.testing.assert_allclose(list(vg.last_error), list(vo.last_error), rtol=1e-6)
    rel = np.abs(xg.vector() - xo.vector()).max() / np.abs(xo.vector()).max()
    assert rel < 1e-9, rel                      # bar (north_star): 1e-5
    np.testing.assert_allclose(xg.P, xo.P, rtol=1e-6, atol=1e-14)
    # idempotence: the same update on the same inputs gives the same bits (deterministic reductions)
    xg2, _, _ = _frame_gpu(flb, h, f)
    assert (bits(xg2.vector()) == bits(xg.vector())).all() and (bits(xg2.P) == bits(xg.P)).all()
    # covariance stays symmetric and shrinks on the observed block
    P = xg.P
    np.testing.assert_allclose(P, P.T, atol=1e-10 * np.abs(P).max())
    assert np.all(np.diag(P)[:6] <= np.diag(f["cov"])[:6] * (1 + 1e-12))
    h.close()


def test_full_size_knn_and_pass_properties(flb, po):
    f = flb.synth.make_frame("C2")
    h = flb.Handle(device=0, cell_size=f["cfg"].cell_size)
    h.load_frame(f)
    prm = flb.capi.lio_params(f, 2)
    g = h.lio_pass(prm, f["R_prop"], f["p_prop"], True, width=12)
    ok = g["nn_idx"][:, 4] >= 0
    assert ok.mean() > 0.95
    d2 = g["nn_d2"][ok]
    assert (np.diff(d2, axis=1) >= 0).all() and (d2[:, 4] <= 5.0).all()          # sorted ascending, bounded
    # distances recomputed from the returned indices in float32, reference op order -> identical bits
    q = g["world"][ok][:, None, :]
    nb = f["map_xyz"][g["nn_idx"][ok]]
    d = (q - nb).astype(np.float32)
    rec = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    assert (bits(rec) == bits(d2)).all()
    # exactness against the reference's own ikd-Tree on the full scan
    if po.ref_lib() is not None:
        ri, rd = po.IkdTreeRef(f["map_xyz"]).knn(g["world"], nthreads=8)
        okr = rd[:, 4] <= 5.0
        assert (okr == ok).all()
        assert (g["nn_idx"][ok] == ri[ok]).all() and (bits(g["nn_d2"][ok]) == bits(rd[ok])).all()
    # normal equations: symmetric, PSD, and consistent with the exported rows
    H = g["rows"]
    np.testing.assert_allclose(g["HTH"], H.T @ H, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(g["HTh"], H.T @ g["meas"], rtol=1e-9, atol=1e-9)
    assert np.linalg.eigvalsh(g["HTH"]).min() > -1e-6
    # 6-wide layout is the permuted leading block of the 12-wide one
    g6 = h.lio_pass(prm, f["R_prop"], f["p_prop"], False, width=6)
    perm = [3, 4, 5, 0, 1, 2]
    np.testing.assert_allclose(g["HTH"][:6, :6][np.ix_(perm, perm)], g6["HTH"], rtol=1e-10)
    h.close()


def test_vio_pass_linearity_full_size(flb, po):
    """HTz is linear in the reference patches: z = I - P, so shifting every reference value by c shifts z by -c."""
    f = flb.synth.make_frame("C2")
    h = flb.Handle(device=0, cell_size=f["cfg"].cell_size)
    h.load_frame(f)
    prm = flb.capi.vio_params(f, 3)
    a = h.vio_pass(prm, f["R_prop"], f["p_prop"], 1)
    h.patches_upload(f["patch_pos"], f["patch_ref"] + np.float32(8.0), f["patch_level"])
    b = h.vio_pass(prm, f["R_prop"], f["p_prop"], 1)
    assert a["n_meas"] == b["n_meas"] == (len(f["patch_pos"]) - a["skipped"]) * 64
    np.testing.assert_allclose(b["z"], a["z"] - 8.0 * (np.abs(a["H_sub"]).sum(1) > 0), atol=1e-4)
    assert (bits(a["H_sub"]) == bits(b["H_sub"])).all()          # the Jacobian does not depend on the reference patch
    np.testing.assert_allclose(a["HTH6"], b["HTH6"], rtol=0, atol=0)
    h.close()

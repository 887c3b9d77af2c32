"""GPU tier: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded synthetic frames.  Bars (BASELINE.json north_star): kNN indices bit-exact; float32
stages (world points, plane, pd2, photometric residuals) bit-exact; double rows / normal
equations / converged state within 1e-5 relative (we assert much tighter)."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu

STATE_RTOL = 1e-9     # bar is 1e-5 (north_star); measured ~1e-12


@pytest.fixture(scope="module", params=[1, 0], ids=["persistent", "kernel-per-pass"])
def handle(flb, request):
    """Both execution modes of the update: one cooperative persistent kernel per update (default)
    and the kernel-per-pass path; they must agree with the oracle (and hence with each other)."""
    h = flb.Handle(device=0, persistent=request.param)
    yield h
    h.close()


def _ostate(po, f):
    return po.state_from_frame(f)


def _gstate(flb, f):
    return flb.capi.State18.from_frame(f)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", ["T0", "T1"])
def test_knn_bit_exact(flb, po, frames, handle, name):
    f = frames(name)
    handle.map_upload(f["map_xyz"])
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    o = lio.run_pass(po.lio_params(f, 3), f["R_prop"], f["p_prop"], True)
    idx, d2 = handle.knn(o["world"])
    ok = o["nn_d2"][:, 4] <= 5.0
    assert (np.diff(o["nn_d2"][ok], axis=1) > 0).all()          # no ties in the fixture
    assert (idx[ok] == o["nn_idx"][ok]).all()
    assert (bits(d2[ok]) == bits(o["nn_d2"][ok])).all()
    assert (idx[~ok][:, 4] == -1).all()


def test_knn_vs_reference_ikdtree(flb, po, frames, handle):
    """Same neighbours as the reference's own KD_TREE::Nearest_Search (oracle/_ref)."""
    if po.ref_lib() is None:
        pytest.skip("oracle/_ref/libikdtree_ref.so not built")
    f = frames("T1")
    handle.map_upload(f["map_xyz"])
    tree = po.IkdTreeRef(f["map_xyz"])
    rng = np.random.default_rng(17)
    q = (f["map_xyz"][rng.integers(0, len(f["map_xyz"]), 5000)] + rng.normal(0, 0.3, (5000, 3))).astype(np.float32)
    ri, rd = tree.knn(q)
    gi, gd = handle.knn(q)
    ok = rd[:, 4] <= 5.0
    assert ok.sum() > 4000
    assert (gi[ok] == ri[ok]).all() and (bits(gd[ok]) == bits(rd[ok])).all()


def test_knn_exact_ties_vs_reference_ikdtree(flb, po, handle):
    """A fixture WITH exact float ties (map on a 0.25 m integer lattice, queries on lattice points and cell
    centres).  ikd-Tree keeps a candidate only if dist < top (strict, ikd_Tree.cpp:860), so among points at exactly
    the same distance the winner is whichever its traversal visits first -- a property of the tree's build, not of
    the data (ikd_Tree.h:57-60 only orders the heap).  What IS defined, and asserted: the five distances agree bit
    for bit at every rank, every returned index is a true neighbour at exactly that distance, and the two index
    lists differ only inside groups of exactly tied distances."""
    if po.ref_lib() is None:
        pytest.skip("oracle/_ref/libikdtree_ref.so not built")
    g = np.arange(-8, 9, dtype=np.float32) * 0.25
    lat = np.stack(np.meshgrid(g, g, g[:5], indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(3)
    lat = lat[rng.permutation(len(lat))]
    handle.map_upload(lat)
    tree = po.IkdTreeRef(lat)
    q = np.concatenate([lat[:400], lat[400:800] + np.float32(0.125), lat[800:1000] + np.array([0.125, 0, 0], np.float32)]).astype(np.float32)
    ri, rd = tree.knn(q)
    gi, gd = handle.knn(q)
    assert (bits(gd) == bits(rd)).all()                        # distances: bit-exact at every rank
    tied_queries = 0
    for k in range(len(q)):
        d = lat[gi[k]] - q[k]
        dd = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        assert (bits(dd.astype(np.float32)) == bits(gd[k])).all()      # every index is a neighbour at that distance
        assert len(set(gi[k].tolist())) == 5
        diff = np.nonzero(gi[k] != ri[k])[0]
        for j in diff:
            assert (rd[k] == rd[k][j]).sum() > 1 or rd[k][j] == rd[k][4]   # only inside an exact tie / at the boundary tie
        tied_queries += len(np.unique(rd[k])) < 5
    assert tied_queries > 900                                   # the fixture really is made of ties


@pytest.mark.parametrize("name", ["T0", "T1"])
@pytest.mark.parametrize("width", [6, 12])
def test_lio_pass_parity(flb, po, frames, handle, name, width):
    f = frames(name)
    handle.load_frame(f)
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    oprm, gprm = po.lio_params(f, 3), flb.capi.lio_params(f, 3)
    # pass 1: rematch at the prior pose; pass 2: cached planes at a perturbed pose
    poses = [(f["R_prop"], f["p_prop"], True),
             (f["R_prop"] @ flb.synth.exp_so3(np.array([0.002, -0.001, 0.003])), f["p_prop"] + [0.01, -0.02, 0.005], False)]
    for R, p, rematch in poses:
        o = lio.run_pass(oprm, R, p, rematch, rows12=True)
        g = handle.lio_pass(gprm, R, p, rematch, width=width)
        assert (bits(g["world"]) == bits(o["world"])).all()
        if rematch:
            ok = o["nn_d2"][:, 4] <= 5.0
            assert (g["nn_idx"][ok] == o["nn_idx"][ok]).all()
            assert (bits(g["nn_d2"][ok]) == bits(o["nn_d2"][ok])).all()
        assert g["n"] == o["n"] and (g["sel_idx"] == o["sel_idx"]).all()
        mask = np.zeros(len(g["rowmask"]), bool)
        mask[o["sel_idx"]] = True
        assert (g["rowmask"].astype(bool) == mask).all()          # the rows in the sums are exactly the oracle's
        sel = o["sel_idx"]
        assert (bits(g["pabcd"][sel]) == bits(o["pabcd"][sel])).all()
        assert (bits(g["pd2"][sel]) == bits(o["pd2"][sel])).all()
        orow = o["Hsub"] if width == 6 else o["h_x"]
        np.testing.assert_allclose(g["rows"], orow, rtol=1e-12, atol=0)
        np.testing.assert_allclose(g["meas"], o["meas"], rtol=0, atol=0)
        oH = o["HTH6"] if width == 6 else o["HTH12"]
        oh = o["HTz6"] if width == 6 else o["HTh12"]
        assert rel(g["HTH"], oH) < 1e-11 and rel(g["HTh"], oh) < 1e-11
        assert abs(g["total_residual"] - o["total_residual"]) <= 1e-12 * max(o["total_residual"], 1)


@pytest.mark.parametrize("name,T,early", [("T0", 4, True), ("T1", 4, True), ("T1", 2, False), ("T0", 0, True), ("T1", 10, True)])
def test_lio_update_parity(flb, po, frames, handle, name, T, early):
    f = frames(name)
    handle.load_frame(f)
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    xo, xpo = _ostate(po, f), _ostate(po, f)
    orep = lio.update(po.lio_params(f, T, early_stop=early), xo, xpo)
    xg, xpg = _gstate(flb, f), _gstate(flb, f)
    grep = handle.lio_update(flb.capi.lio_params(f, T, early_stop=early), xg, xpg)
    assert (grep.passes, grep.knn_passes, grep.n_eff_last) == (orep.passes, orep.knn_passes, orep.n_eff_last)
    assert grep.rows_total == orep.rows_total and grep.converged_last == orep.converged_last
    assert rel(xg.vector(), xo.vector()) < STATE_RTOL
    assert rel(xg.P, xo.P) < 1e-7
    assert abs(grep.res_mean_last - orep.res_mean_last) < 1e-9


@pytest.mark.parametrize("name", ["T0", "T1"])
@pytest.mark.parametrize("level", [2, 1, 0])
def test_vio_pass_parity(flb, po, frames, handle, name, level):
    f = frames(name)
    handle.load_frame(f)
    vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
    o = vio.run_pass(po.vio_params(f, 3), f["R_prop"], f["p_prop"], level)
    g = handle.vio_pass(flb.capi.vio_params(f, 3), f["R_prop"], f["p_prop"], level)
    assert g["n_meas"] == o["n_meas"] and g["skipped"] == o["skipped"]
    assert (bits(g["z"]) == bits(o["z"])).all()
    assert (bits(g["errors"]) == bits(o["errors"])).all()
    assert bits(np.float32(g["error"])) == bits(np.float32(o["error"]))
    np.testing.assert_allclose(g["H_sub"], o["H_sub"], rtol=1e-12, atol=1e-300)
    assert rel(g["HTH6"], o["HTH6"]) < 1e-11 and rel(g["HTz6"], o["HTz6"]) < 1e-10


@pytest.mark.parametrize("name,T,early,force", [("T0", 4, True, False), ("T1", 4, True, False), ("T1", 3, False, True),
                                                ("T0", 10, True, False)])
def test_vio_update_parity(flb, po, frames, handle, name, T, early, force):
    f = frames(name)
    handle.load_frame(f)
    vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
    xo, xpo = _ostate(po, f), _ostate(po, f)
    orep = vio.update(po.vio_params(f, T, early_stop=early, force_all_passes=force), xo, xpo)
    xg, xpg = _gstate(flb, f), _gstate(flb, f)
    grep = handle.vio_update(flb.capi.vio_params(f, T, early_stop=early, force_all_passes=force), xg, xpg)
    assert list(grep.passes) == list(orep.passes)
    assert grep.rows_total == orep.rows_total and grep.cov_updated == orep.cov_updated
    np.testing.assert_allclose(list(grep.last_error), list(orep.last_error), rtol=1e-6)
    assert rel(xg.vector(), xo.vector()) < STATE_RTOL
    assert rel(xg.P, xo.P) < 1e-7
    # the early-stop scenarios contain rejected steps (error > last_error, :888-892): the persistent kernel publishes
    # the accept branch ahead of the error sum, so these are the cases where it has to discard a pass and roll back
    if early and (name, T) in (("T1", 4), ("T0", 10)):
        assert orep.rejects >= 1
    # sub_sparse_map->errors as the last executed pass left them (:851), bit for bit
    assert (bits(handle.vio_errors()) == bits(vio.errors())).all()


def test_frame_chain_parity(flb, po, frames, handle):
    """LIO update then VIO update with the LIO posterior as the VIO prior (bench 'frame')."""
    f = frames("T1")
    handle.load_frame(f)
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
    xo, xpo = _ostate(po, f), _ostate(po, f)
    lio.update(po.lio_params(f, 3), xo, xpo)
    xpo2 = xo.copy()
    vio.update(po.vio_params(f, 3), xo, xpo2)
    xg, xpg = _gstate(flb, f), _gstate(flb, f)
    handle.lio_update(flb.capi.lio_params(f, 3), xg, xpg)
    xpg2 = xg.copy()
    handle.vio_update(flb.capi.vio_params(f, 3), xg, xpg2)
    assert rel(xg.vector(), xo.vector()) < STATE_RTOL
    # the update must actually improve on the prior
    assert np.linalg.norm(xg.p - f["p_true"]) < np.linalg.norm(f["p_prop"] - f["p_true"])


def test_call_order_and_argument_errors(flb):
    h = flb.Handle(device=0)
    f = flb.synth.make_frame("T0")
    prm = flb.capi.lio_params(f, 3)
    x = flb.capi.State18.from_frame(f)
    with pytest.raises(flb.FlbError) as e:
        h.lio_update(prm, x, x.copy())
    assert e.value.code == -4
    with pytest.raises(flb.FlbError) as e:
        h.knn(np.zeros((4, 3), np.float32))
    assert e.value.code == -4
    with pytest.raises(flb.FlbError) as e:
        h.map_upload(np.full((8, 3), np.nan, np.float32))
    assert e.value.code == -1
    h.close()


def test_gpu_matches_committed_golden(flb, handle):
    """The CUDA path against tests/golden/T0_golden.npz (no generator, no oracle at run time)."""
    from golden_util import load_golden
    frame, g = load_golden()
    handle.load_frame(frame)
    o = handle.lio_pass(flb.capi.lio_params(frame, 3), frame["R_prop"], frame["p_prop"], True, width=12)
    assert (bits(o["world"]) == bits(g["lio_world"])).all()
    ok = g["lio_nn_d2"][:, 4] <= 5.0
    assert (o["nn_idx"][ok] == g["lio_nn_idx"][ok]).all()
    sel = g["lio_sel_idx"]
    assert (o["sel_idx"] == sel).all()
    assert (bits(o["pabcd"][sel]) == bits(g["lio_pabcd"][sel])).all()
    assert (bits(o["pd2"][sel]) == bits(g["lio_pd2"][sel])).all()
    np.testing.assert_allclose(o["rows"], g["lio_h_x"], rtol=1e-12)
    assert rel(o["HTH"], g["lio_HTH12"]) < 1e-11
    x = flb.capi.State18.from_frame(frame)
    rep = handle.lio_update(flb.capi.lio_params(frame, 4), x, x.copy())
    assert [rep.passes, rep.knn_passes, rep.n_eff_last, rep.rows_total] == list(g["lio_report"])
    assert rel(x.vector(), g["lio_state"]) < STATE_RTOL
    for level in (2, 0):
        v = handle.vio_pass(flb.capi.vio_params(frame, 3), frame["R_prop"], frame["p_prop"], level)
        assert (bits(v["z"]) == bits(g[f"vio{level}_z"])).all()
        assert (bits(v["errors"]) == bits(g[f"vio{level}_errors"])).all()
    xv = x.copy()
    vrep = handle.vio_update(flb.capi.vio_params(frame, 4), xv, x.copy())
    assert [*vrep.passes, vrep.rows_total, vrep.cov_updated] == list(g["vio_report"])
    assert rel(xv.vector(), g["vio_state"]) < STATE_RTOL


@pytest.mark.parametrize("name,T", [("T0", 10), ("T1", 4)])
def test_vio_update_level_parity(flb, po, frames, name, T):
    """flb_vio_update_level == LidarSelector::UpdateState alone: driving the three levels from the host with the
    reference's own ComputeJ loop (:974-981) reproduces flb_vio_update and the oracle, level by level."""
    f = frames(name)
    h = flb.Handle(device=0)
    h.load_frame(f)
    vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
    oprm, gprm = po.vio_params(f, T), flb.capi.vio_params(f, T)
    xo, xg = _ostate(po, f), _gstate(flb, f)
    xpo, xpg = xo.copy(), xg.copy()
    Go = np.zeros((18, 18))
    Gg = np.zeros((18, 18))
    now_o = now_g = np.float32(1e10)
    for level in (2, 1, 0):
        now_o, Go, ro = vio.update_level(oprm, level, 1e10, xo, xpo, Go)
        now_g, G6, rg = h.vio_update_level(gprm, level, 1e10, xg, xpg)
        if rg.passes[level] > 0 and now_g < 1e10:
            Gg[:, :6] = G6
        assert list(rg.passes) == list(ro.passes)
        assert abs(now_g - now_o) <= 1e-6 * abs(now_o)
        assert rel(xg.vector(), xo.vector()) < STATE_RTOL
        assert np.abs(Gg - Go).max() < 1e-9 * max(np.abs(Go).max(), 1e-30)
        assert (bits(h.vio_errors()) == bits(vio.errors())).all()
    # ... and the whole of ComputeJ assembled from the three calls equals flb_vio_update
    Pg = xg.P - Gg @ xg.P
    xw = _gstate(flb, f)
    h.vio_update(gprm, xw, xw.copy())
    assert rel(xg.vector(), xw.vector()) < 1e-12 and rel(Pg, xw.P) < 1e-9
    h.close()

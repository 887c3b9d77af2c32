"""CPU tier: the product's per-thread device math (flb_device.cuh compiled for the host by
tests/hostemu) against the oracle.  Float32 stages must be BIT-EXACT."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def grid_build(map_xyz, cell, max_d2=5.0):
    """numpy mirror of flb_map_upload's grid geometry + k_map_cell_ids + the stable sort."""
    m = np.ascontiguousarray(map_xyz, np.float32)
    lo, hi = m.min(0), m.max(0)
    dims = (np.floor((hi.astype(np.float64) - lo) / cell) + 1).astype(np.int64)
    cellf = np.float32(cell)
    inv = np.float32(1.0) / cellf
    c = np.floor((m - lo) * inv).astype(np.int64)
    c = np.clip(c, 0, dims - 1)
    key = (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]
    order = np.argsort(key, kind="stable")
    ncell = int(np.prod(dims))
    cell_start = np.searchsorted(key[order], np.arange(ncell + 1), side="left").astype(np.int32)
    pts4 = np.zeros((len(m), 4), np.float32)
    pts4[:, :3] = m[order]
    pts4[:, 3] = order.astype(np.int32).view(np.float32)
    gridf = np.array([lo[0], lo[1], lo[2], cellf, inv, max_d2], np.float32)
    gridi = np.array([dims[0], dims[1], dims[2], int(np.ceil(np.sqrt(max_d2) / cell)) + 1], np.int32)
    return gridf, gridi, cell_start, pts4, order


@pytest.mark.parametrize("name", ["T0", "T1"])
def test_knn_grid_matches_bruteforce(hostemu, po, frames, name):
    f = frames(name)
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    o = lio.run_pass(po.lio_params(f, 3), f["R_prop"], f["p_prop"], True)
    q = o["world"]
    gridf, gridi, cell_start, pts4, order = grid_build(f["map_xyz"], f["cfg"].cell_size)
    pos = np.empty((len(q), 5), np.int32)
    d2 = np.empty((len(q), 5), np.float32)
    hostemu.emu_knn(_p(gridf), _p(gridi), _p(cell_start), _p(pts4), _p(q), len(q), _p(pos), _p(d2))
    idx = np.where(pos >= 0, order[np.maximum(pos, 0)], -1)
    ok = o["nn_d2"][:, 4] <= 5.0          # the reference rejects the rest (src/laserMapping.cpp:1549)
    assert ok.sum() > 0.9 * len(q)
    # no exact distance ties in the fixture, then demand exact equality
    assert (np.diff(o["nn_d2"][ok], axis=1) > 0).all()
    assert (idx[ok] == o["nn_idx"][ok]).all()
    assert (bits(d2[ok]) == bits(o["nn_d2"][ok])).all()
    # rejected queries must also be rejected by the bounded search
    assert (~(d2[~ok][:, 4] <= 5.0)).all()


@pytest.mark.parametrize("cell", [0.6, 0.25])
def test_knn_grid_far_and_outside_queries(hostemu, po, frames, cell):
    """Queries anywhere in (and around) the map volume: most are far from every surface, so the walk goes through the
    ring >= 2 shells (nine-run batches, grid clipping) up to the search radius -- exact against brute force."""
    f = frames("T0")
    rng = np.random.default_rng(5)
    lo, hi = f["map_xyz"].min(0), f["map_xyz"].max(0)
    q = rng.uniform(lo - 6, hi + 6, size=(4000 if cell > 0.5 else 1500, 3)).astype(np.float32)
    bi, bd = po.knn_brute(f["map_xyz"], q)
    gridf, gridi, cell_start, pts4, order = grid_build(f["map_xyz"], cell)
    pos = np.empty((len(q), 5), np.int32)
    d2 = np.empty((len(q), 5), np.float32)
    hostemu.emu_knn(_p(gridf), _p(gridi), _p(cell_start), _p(pts4), _p(q), len(q), _p(pos), _p(d2))
    idx = np.where(pos >= 0, order[np.maximum(pos, 0)], -1)
    ok = bd[:, 4] <= 5.0
    assert ok.any() and (~ok).any()
    assert (idx[ok] == bi[ok]).all() and (bits(d2[ok]) == bits(bd[ok])).all()
    assert (~(d2[~ok][:, 4] <= 5.0)).all()
    # partial results (k-th neighbour within range) are exact prefix-wise
    for j in range(5):
        okj = bd[:, j] <= 5.0
        assert (idx[okj, j] == bi[okj, j]).all()


@pytest.mark.parametrize("cell", [0.3, 0.45, 1.1, 2.5])
def test_knn_grid_any_cell_size(hostemu, po, frames, cell):
    f = frames("T0")
    rng = np.random.default_rng(11)
    q = (f["map_xyz"][rng.integers(0, len(f["map_xyz"]), 1500)] + rng.normal(0, 0.2, (1500, 3))).astype(np.float32)
    bi, bd = po.knn_brute(f["map_xyz"], q)
    gridf, gridi, cell_start, pts4, order = grid_build(f["map_xyz"], cell)
    pos = np.empty((len(q), 5), np.int32)
    d2 = np.empty((len(q), 5), np.float32)
    hostemu.emu_knn(_p(gridf), _p(gridi), _p(cell_start), _p(pts4), _p(q), len(q), _p(pos), _p(d2))
    idx = order[np.maximum(pos, 0)]
    assert (idx == bi).all() and (bits(d2) == bits(bd)).all()


def test_plane_fit_bit_exact(hostemu, po, frames):
    f = frames("T1")
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    o = lio.run_pass(po.lio_params(f, 3), f["R_prop"], f["p_prop"], True)
    rng = np.random.default_rng(3)
    n_ok = n_bad = 0
    for i in rng.choice(len(o["nn_idx"]), 1200, replace=False):
        if o["nn_idx"][i, 4] < 0:
            continue
        nb = f["map_xyz"][o["nn_idx"][i]].astype(np.float32)
        ok_o, pl_o = po.esti_plane(nb)
        out = np.zeros(4, np.float32)
        ok_e = hostemu.emu_plane_fit(_p(np.ascontiguousarray(nb)), C.c_float(0.1), _p(out))
        assert bool(ok_e) == ok_o
        assert (bits(out) == bits(pl_o)).all()
        n_ok += ok_o
        n_bad += not ok_o
    assert n_ok > 500
    # degenerate / far-field / rank-deficient inputs
    cases = [np.zeros((5, 3), np.float32),
             np.tile(np.array([[1.0, 2.0, 3.0]], np.float32), (5, 1)),
             (np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [4, 0, 0]], np.float32) + 100.0),
             (rng.normal(0, 0.2, (5, 3)) * [1, 1, 0.001] + [90.0, -80.0, 3.0]).astype(np.float32)]
    for nb in cases:
        ok_o, pl_o = po.esti_plane(nb)
        out = np.zeros(4, np.float32)
        ok_e = hostemu.emu_plane_fit(_p(np.ascontiguousarray(nb)), C.c_float(0.1), _p(out))
        assert bool(ok_e) == ok_o
        assert (bits(out) == bits(pl_o)).all() or (np.isnan(out) == np.isnan(pl_o)).all()


def test_lio_point_math_bit_exact(hostemu, po, frames):
    f = frames("T1")
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    prm = po.lio_params(f, 3)
    o = lio.run_pass(prm, f["R_prop"], f["p_prop"], True, rows12=True)
    N = lio.N
    R = np.ascontiguousarray(f["R_prop"], np.float64)
    p = np.ascontiguousarray(f["p_prop"], np.float64)
    RLI = np.ascontiguousarray(f["R_LI"], np.float64)
    tLI = np.ascontiguousarray(f["t_LI"], np.float64)
    body = np.ascontiguousarray(f["scan_body"], np.float32)
    world = np.empty((N, 3), np.float32)
    pd2 = np.empty(N, np.float32)
    gate = np.empty(N, np.uint8)
    row6 = np.empty((N, 6))
    row12 = np.empty((N, 12))
    pabcd = np.ascontiguousarray(o["pabcd"])
    hostemu.emu_lio_points(_p(R), _p(p), _p(RLI), _p(tLI), _p(body), N, _p(pabcd), _p(world), _p(pd2), _p(gate), _p(row6),
                           _p(row12))
    assert (bits(world) == bits(o["world"])).all()
    sel = o["sel_idx"]
    assert len(sel) > 0.8 * N
    assert (bits(pd2[sel]) == bits(o["pd2"][sel])).all()
    assert gate[sel].all()
    assert (bits(row6[sel]) == bits(o["Hsub"])).all()
    assert (bits(row12[sel]) == bits(o["h_x"])).all()


@pytest.mark.parametrize("name,level", [("T0", 2), ("T0", 0), ("T1", 1), ("T1", 2)])
def test_vio_pixel_math_bit_exact(hostemu, po, frames, name, level):
    f = frames(name)
    vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
    prm = po.vio_params(f, 3)
    o = vio.run_pass(prm, f["R_prop"], f["p_prop"], level)
    cam = f["cam"]
    camv = np.array([cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], *cam["d"]], np.float64)
    Rli = f["R_LI"].T
    Pli = -f["R_LI"].T @ f["t_LI"]
    Rci = np.ascontiguousarray(f["Rcl"] @ Rli)
    Pci = np.ascontiguousarray(f["Rcl"] @ Pli + f["Pcl"])
    Pn = vio.Pn
    z = np.empty(Pn * 64)
    H = np.empty((Pn * 64, 6))
    err = np.empty(Pn, np.float32)
    valid = np.empty(Pn, np.uint8)
    R = np.ascontiguousarray(f["R_prop"], np.float64)
    p = np.ascontiguousarray(f["p_prop"], np.float64)
    img = np.ascontiguousarray(f["image"])
    hostemu.emu_vio(_p(camv), _p(Rci), _p(Pci), _p(R), _p(p), _p(vio.pos), _p(vio.patch), _p(vio.level), Pn, level, _p(img),
                    _p(z), _p(H), _p(err), _p(valid))
    assert Pn - valid.sum() == o["skipped"]
    assert valid.sum() > 0.8 * Pn
    assert (bits(z) == bits(o["z"])).all()                 # float32 taps -> bit-exact residuals
    assert (bits(err) == bits(o["errors"])).all()
    np.testing.assert_allclose(H, o["H_sub"], rtol=1e-12, atol=1e-300)
    # Rci / Pci are formed with numpy here (different op order than the oracle's init()), so
    # Jacobians are compared to 1e-12; with identical constants they agree to the last bit.


def test_so3_matches_oracle(hostemu, po):
    import ctypes as C
    rng = np.random.default_rng(0)
    L = po.lib()
    for s in (1e-9, 1e-6, 1e-4, 1e-2, 0.5, 3.0):
        v = rng.normal(size=3) * s
        Ro, Re = np.empty(9), np.empty(9)
        L.flo_exp3(_p(v), _p(Ro))
        hostemu.emu_exp3(_p(v), _p(Re))
        assert (bits(Ro) == bits(Re)).all()
        lo, le = np.empty(3), np.empty(3)
        L.flo_log3(_p(Ro), _p(lo))
        hostemu.emu_log3(_p(Ro), _p(le))
        assert (bits(lo) == bits(le)).all()

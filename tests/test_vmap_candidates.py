"""Row f4: candidate scoring of LidarSelector::addSparseMap's first loop
(reference src/lidar_selection.cpp:150-168; vikit shiTomasiScore / isInFrame / world2cam restated).
CPU tier: the product's per-point device math compiled for the host (tests/hostemu) against the oracle, bit-exact.
GPU tier: flb_visual_candidates (atomic per-cell competition) against the oracle's sequential loop, bit-exact."""
import ctypes as C

import numpy as np
import pytest

GRID, BORDER = 40, 40


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _case(flb, name, seed, n=6000, behind=True):
    f = flb.synth.make_frame(name)
    rng = np.random.default_rng(seed)
    Rcw = f["Rcl"] @ f["R_LI"].T @ f["R_prop"].T
    Pcw = -Rcw @ f["p_prop"] + f["Rcl"] @ (-f["R_LI"].T @ f["t_LI"]) + f["Pcl"]
    pts = ((f["R_prop"] @ (f["R_LI"] @ f["scan_body"].T.astype(np.float64) + f["t_LI"][:, None])).T + f["p_prop"]).astype(np.float32)
    pts = pts[rng.choice(len(pts), min(n, len(pts)), replace=False)]
    if behind:        # a few points behind / beside the camera and duplicates (equal scores -> first index wins)
        extra = rng.normal(0, 6.0, (200, 3)).astype(np.float32)
        pts = np.concatenate([pts, extra, pts[:150]])
    cam = f["cam"]
    ncell = (cam["width"] // GRID) * (cam["height"] // GRID)
    seed_vals = np.where(rng.random(ncell) < 0.3, rng.uniform(0, 400, ncell), 0.0).astype(np.float32)
    return f, Rcw, Pcw, pts, seed_vals


def test_shi_tomasi_hostemu_bit_exact(flb, po, hostemu):
    f = flb.synth.make_frame("T1")
    img = np.ascontiguousarray(f["image"], np.uint8)
    h, w = img.shape
    rng = np.random.default_rng(2)
    hostemu.emu_shi_tomasi.restype = C.c_float
    for u, v in np.concatenate([rng.integers(-3, [w + 3, h + 3], (3000, 2)), [[4, 4], [5, 5], [w - 6, h - 6], [w - 5, h - 5]]]):
        a = np.float32(hostemu.emu_shi_tomasi(_p(img), w, h, int(u), int(v)))
        b = po.shi_tomasi(img, u, v)
        assert a.view(np.uint32) == b.view(np.uint32), (u, v, a, b)
    flat = np.full((64, 64), 77, np.uint8)
    assert po.shi_tomasi(flat, 30, 30) == 0.0
    corner = flat.copy(); corner[32:, 32:] = 200
    assert po.shi_tomasi(corner, 32, 32) > po.shi_tomasi(corner, 32, 10) >= 0.0      # corner beats edge


@pytest.mark.parametrize("name", ["T0", "T1"])
def test_candidate_math_hostemu_bit_exact(flb, po, hostemu, name):
    f, Rcw, Pcw, pts, seed_vals = _case(flb, name, 5)
    cam = f["cam"]
    camv = np.array([cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], *cam["d"]], np.float64)
    img = np.ascontiguousarray(f["image"], np.uint8)
    cell = np.zeros(len(pts), np.int32); score = np.zeros(len(pts), np.float32)
    hostemu.emu_visual_candidates(_p(camv), _p(np.ascontiguousarray(Rcw)), _p(np.ascontiguousarray(Pcw)), _p(img), _p(pts),
                                  len(pts), GRID, BORDER, _p(cell), _p(score))
    # replay the reference's sequential competition on the per-point results and compare with the oracle's loop
    mv = seed_vals.copy(); win = np.full(len(mv), -1, np.int32)
    for i in np.nonzero(cell >= 0)[0]:
        if score[i] > mv[cell[i]]:
            mv[cell[i]] = score[i]; win[cell[i]] = i
    mv_o, win_o = po.visual_candidates(cam, Rcw, Pcw, img, pts, GRID, BORDER, seed_vals)
    assert (cell >= 0).sum() > 100 and (win_o >= 0).sum() > 10
    assert (win == win_o).all() and (mv.view(np.uint32) == mv_o.view(np.uint32)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("T0", 2000), ("T1", 6000), ("C2", 24000)])
def test_visual_candidates_gpu_bit_exact(flb, po, name, n):
    f, Rcw, Pcw, pts, seed_vals = _case(flb, name, 9, n=n)
    h = flb.Handle(device=0)
    h.camera_set(f["cam"]); h.image_upload(f["image"])
    mv_o, win_o = po.visual_candidates(f["cam"], Rcw, Pcw, f["image"], pts, GRID, BORDER, seed_vals)
    mv, win = h.visual_candidates(Rcw, Pcw, pts, GRID, BORDER, seed_vals)
    assert (win == win_o).all() and (mv.view(np.uint32) == mv_o.view(np.uint32)).all()
    # no points: nothing changes; a second call with the updated values changes nothing either (strict >)
    mv0, win0 = h.visual_candidates(Rcw, Pcw, pts[:0], GRID, BORDER, seed_vals)
    assert (win0 == -1).all() and (mv0 == seed_vals).all()
    mv2, win2 = h.visual_candidates(Rcw, Pcw, pts, GRID, BORDER, mv)
    assert (win2 == -1).all() and (mv2.view(np.uint32) == mv.view(np.uint32)).all()
    # the point with index 0 can win its cell (round 1 encoded "no winner" and "point 0" alike): put a winner first
    k = int(win_o[win_o >= 0][0])
    order = np.concatenate([[k], np.delete(np.arange(len(pts)), k)])
    mv_o3, win_o3 = po.visual_candidates(f["cam"], Rcw, Pcw, f["image"], pts[order], GRID, BORDER, seed_vals)
    mv3, win3 = h.visual_candidates(Rcw, Pcw, pts[order], GRID, BORDER, seed_vals)
    assert (win_o3 == 0).sum() == 1 and (win3 == win_o3).all() and (mv3.view(np.uint32) == mv_o3.view(np.uint32)).all()
    h.close()

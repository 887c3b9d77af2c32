"""Batched frames: B independent frames advanced by one launch per pass must give, frame
by frame, the bit-identical result of the same frame run alone through the kernel-per-pass path -- and hence the
oracle's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(flb, f, B):
    """B variations of one synthetic frame: different sub-scans and different priors, same map / image / patches."""
    rng = np.random.default_rng(11)
    out = []
    for b in range(B):
        n = len(f["scan_body"]) - 37 * b
        idx = rng.permutation(len(f["scan_body"]))[:n]
        R = f["R_prop"] @ flb.synth.exp_so3(rng.normal(0, 0.002, 3))
        p = f["p_prop"] + rng.normal(0, 0.01, 3)
        x = flb.capi.State18.make(R, p, f["vel"], f["bg"], f["ba"], f["grav"], f["cov"])
        out.append((np.ascontiguousarray(f["scan_body"][idx]), x))
    return out


@pytest.mark.parametrize("name,early", [("T0", True), ("T1", False)])
def test_batched_frames_equal_single_frames(flb, po, frames, name, early):
    f = frames(name)
    B = 4
    fr = _frames(flb, f, B)
    lprm = flb.capi.lio_params(f, 3, early_stop=early)
    vprm = flb.capi.vio_params(f, 3, early_stop=early, force_all_passes=not early)
    # reference: each frame alone, kernel-per-pass path
    single = []
    h1 = flb.Handle(device=0, persistent=0)
    h1.load_frame(f)
    for scan, x in fr:
        h1.scan_upload(scan)
        h1.state_upload(x, x.copy())
        h1.lio_update_enqueue(lprm)
        h1.state_set_prior_enqueue()
        h1.vio_update_enqueue(vprm)
        single.append(h1.state_download())
    h1.close()
    hb = flb.Handle(device=0)
    hb.load_frame(f)
    hb.batch_begin(B, max(len(s) for s, _ in fr))
    for b, (scan, x) in enumerate(fr):
        hb.batch_set_frame(b, scan, x, x.copy())
    for rep in range(2):                      # the second round starts from the restored priors
        hb.batch_state_reset_enqueue()
        hb.batch_update_enqueue(lprm, vprm)
        for b in range(B):
            xb, lb, vb = hb.batch_state_download(b)
            xs, ls, vs = single[b]
            assert (lb.passes, lb.knn_passes, lb.n_eff_last, lb.rows_total) == (ls.passes, ls.knn_passes, ls.n_eff_last, ls.rows_total)
            assert list(vb.passes) == list(vs.passes) and vb.rows_total == vs.rows_total and vb.cov_updated == vs.cov_updated
            assert (xb.vector() == xs.vector()).all() and (np.array(xb.cov[:]) == np.array(xs.cov[:])).all()
    hb.close()
    # and one of them against the oracle
    scan, x = fr[1]
    lio = po.Lio(f["map_xyz"], scan)
    vio = po.Vio(f["image"], f["patch_pos"], f["patch_ref"], f["patch_level"], f["cam"])
    xo = po.State18.make(np.array(x.rot[:]).reshape(3, 3), np.array(x.pos[:]), f["vel"], f["bg"], f["ba"], f["grav"], f["cov"])
    lio.update(po.lio_params(f, 3, early_stop=early), xo, xo.copy())
    vio.update(po.vio_params(f, 3, early_stop=early, force_all_passes=not early), xo, xo.copy())
    assert np.abs(single[1][0].vector() - xo.vector()).max() / np.abs(xo.vector()).max() < 1e-9

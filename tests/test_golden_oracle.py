"""CPU tier: the oracle against the committed golden vectors (tests/golden/T0_golden.npz,
made by tests/golden/make_golden.py).  Float32 stages are bit-exact; double stages to 1e-12."""
import numpy as np

from conftest import bits
from golden_util import load_golden


def test_generator_reproduces_golden_inputs(flb):
    frame, g = load_golden()
    f = flb.synth.make_frame("T0")
    for k in ("map_xyz", "scan_body", "image", "patch_pos", "patch_ref", "patch_level", "R_prop", "p_prop"):
        assert np.array_equal(f[k], frame[k]), f"synthetic generator drifted on {k}"


def test_oracle_lio_pass_matches_golden(po):
    frame, g = load_golden()
    lio = po.Lio(frame["map_xyz"], frame["scan_body"])
    o = lio.run_pass(po.lio_params(frame, 3), frame["R_prop"], frame["p_prop"], True, rows12=True)
    for k in ("world", "nn_d2", "pabcd", "pd2"):
        assert (bits(o[k]) == bits(g["lio_" + k])).all(), k
    assert (o["nn_idx"] == g["lio_nn_idx"]).all() and (o["sel_idx"] == g["lio_sel_idx"]).all()
    for k in ("Hsub", "h_x", "meas", "HTH6", "HTz6", "HTH12", "HTh12"):
        np.testing.assert_allclose(o[k], g["lio_" + k], rtol=1e-12, atol=1e-14)


def test_oracle_updates_match_golden(po):
    frame, g = load_golden()
    lio = po.Lio(frame["map_xyz"], frame["scan_body"])
    x = po.state_from_frame(frame)
    rep = lio.update(po.lio_params(frame, 4), x, x.copy())
    assert [rep.passes, rep.knn_passes, rep.n_eff_last, rep.rows_total] == list(g["lio_report"])
    np.testing.assert_allclose(x.vector(), g["lio_state"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(x.P, g["lio_cov"], rtol=1e-8, atol=1e-14)
    vio = po.Vio(frame["image"], frame["patch_pos"], frame["patch_ref"], frame["patch_level"], frame["cam"])
    for level in (2, 0):
        v = vio.run_pass(po.vio_params(frame, 3), frame["R_prop"], frame["p_prop"], level)
        assert (bits(v["z"]) == bits(g[f"vio{level}_z"])).all()
        assert (bits(v["errors"]) == bits(g[f"vio{level}_errors"])).all()
        np.testing.assert_allclose(v["H_sub"][:8 * 64], g[f"vio{level}_H"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(v["HTH6"], g[f"vio{level}_HTH6"], rtol=1e-11)
    xv = x.copy()
    vrep = vio.update(po.vio_params(frame, 4), xv, x.copy())
    assert [*vrep.passes, vrep.rows_total, vrep.cov_updated] == list(g["vio_report"])
    np.testing.assert_allclose(xv.vector(), g["vio_state"], rtol=1e-10, atol=1e-13)

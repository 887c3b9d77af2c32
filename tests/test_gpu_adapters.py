"""GPU tier: the header-only C++ adapters (Eigen/PCL/OpenCV-shaped call sites -> C ABI), driven through
a small C++ program with mock types (tests/adapters/test_adapters.cpp), checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "adapters", "test_adapters.cpp")
EXE = os.path.join(ROOT, "tests", "adapters", "test_adapters")


def build_exe():
    hdr = os.path.join(ROOT, "fast-livo_b200", "adapters", "fastlivo_b200_adapters.hpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-o", EXE, SRC,
                               "-L" + os.path.join(ROOT, "fast-livo_b200"), "-lfastlivo_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "fast-livo_b200")])


def test_adapters_compile(flb):
    """CPU tier: the adapters compile against mock reference types and link against the C ABI."""
    flb.build()
    build_exe()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_adapters_match_oracle(flb, po, frames, tmp_path):
    flb.build()
    build_exe()
    f = frames("T1")
    T = 3
    cam = f["cam"]
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    Pn = len(f["patch_pos"])
    with open(inp, "wb") as fh:
        fh.write(np.array([len(f["map_xyz"]), len(f["scan_body"]), Pn, cam["width"], cam["height"], T], np.int32).tobytes())
        fh.write(np.ascontiguousarray(f["map_xyz"], np.float32).tobytes())
        fh.write(np.ascontiguousarray(f["scan_body"], np.float32).tobytes())
        fh.write(np.concatenate([f["R_prop"].ravel(), f["p_prop"], f["cov"].ravel(), f["grav"]]).astype(np.float64).tobytes())
        fh.write(np.concatenate([f["R_LI"].ravel(), f["t_LI"], f["Rcl"].ravel(), f["Pcl"]]).astype(np.float64).tobytes())
        fh.write(np.array([cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], *cam["d"]], np.float64).tobytes())
        fh.write(np.array([f["cfg"].laser_point_cov, f["cfg"].img_point_cov], np.float64).tobytes())
        fh.write(np.ascontiguousarray(f["image"]).tobytes())
        fh.write(np.ascontiguousarray(f["patch_pos"], np.float64).tobytes())
        fh.write(np.ascontiguousarray(f["patch_ref"], np.float32).tobytes())
        fh.write(np.ascontiguousarray(f["patch_level"], np.int32).tobytes())
    r = subprocess.run([EXE, str(inp), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    hdr = struct.unpack("8i", raw[:32])
    effct, lpasses, neff, vp0, vp1, vp2, covup, nidx = hdr
    off = 32
    res_mean = struct.unpack("d", raw[off:off + 8])[0]
    off += 8
    h_x = np.frombuffer(raw, np.float64, effct * 12, off).reshape(effct, 12)
    off += effct * 12 * 8
    hvec = np.frombuffer(raw, np.float64, effct, off)
    off += effct * 8
    eidx = np.frombuffer(raw, np.int32, nidx, off)
    off += nidx * 4
    st = np.frombuffer(raw, np.float64, 12 + 324, off)
    off += (12 + 324) * 8
    errs = np.frombuffer(raw, np.float32, Pn, off)

    # oracle: same three steps
    lio = po.Lio(f["map_xyz"], f["scan_body"])
    o = lio.run_pass(po.lio_params(f, T), f["R_prop"], f["p_prop"], True, rows12=True)
    assert effct == o["n"] and (eidx == o["sel_idx"]).all()
    np.testing.assert_allclose(h_x, o["h_x"], rtol=1e-12)
    np.testing.assert_allclose(hvec, o["meas"], rtol=0, atol=0)
    assert abs(res_mean - o["total_residual"] / o["n"]) < 1e-12
    x = po.state_from_frame(f)
    x.vel[:] = [0, 0, 0]
    orep = lio.update(po.lio_params(f, T), x, x.copy())
    assert (lpasses, neff) == (orep.passes, orep.n_eff_last)
    keep = np.array([i % 17 != 5 for i in range(Pn)])           # the C++ side nulls every 17th voxel point
    vio = po.Vio(f["image"], f["patch_pos"][keep], f["patch_ref"][keep], f["patch_level"][keep], f["cam"])
    vrep = vio.update(po.vio_params(f, T), x, x.copy())
    assert [vp0, vp1, vp2] == list(vrep.passes) and covup == vrep.cov_updated
    # sub_sparse_map->errors written back by compute_j (:851): the last pass's per-patch errors; null points keep theirs
    assert (errs[keep].view(np.uint32) == vio.errors().view(np.uint32)).all() and (errs[~keep] == -1.0).all()
    ref = np.concatenate([x.R.ravel(), x.p, x.P.ravel()])
    assert np.abs(st[:12] - ref[:12]).max() / np.abs(ref[:12]).max() < 1e-9
    np.testing.assert_allclose(st[12:], ref[12:], rtol=1e-6, atol=1e-14)


SRC_IMU = os.path.join(ROOT, "tests", "adapters", "test_adapter_imu.cpp")
EXE_IMU = os.path.join(ROOT, "tests", "adapters", "test_adapter_imu")


def build_exe_imu():
    hdr = os.path.join(ROOT, "fast-livo_b200", "adapters", "fastlivo_b200_adapters.hpp")
    if not os.path.exists(EXE_IMU) or os.path.getmtime(EXE_IMU) < max(os.path.getmtime(SRC_IMU), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-o", EXE_IMU, SRC_IMU,
                               "-L" + os.path.join(ROOT, "fast-livo_b200"), "-lfastlivo_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "fast-livo_b200")])


def test_imu_adapter_compiles(flb):
    flb.build()
    build_exe_imu()
    assert os.path.exists(EXE_IMU)


@pytest.mark.gpu
def test_imu_adapter_matches_oracle(flb, po, tmp_path):
    """flb::undistort_pcl with mock sensor_msgs::Imu / PCL / StatesGroup types == the oracle's UndistortPcl."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from imu_util import oracle_inputs
    flb.build()
    build_exe_imu()
    f = flb.synth.make_imu_frame(seed=31, n_points=5000)
    inp, out = tmp_path / "imu_in.bin", tmp_path / "imu_out.bin"
    n = len(f["pts"])
    with open(inp, "wb") as fh:
        fh.write(np.array([len(f["v_imu"]), n], np.int32).tobytes())
        fh.write(np.ascontiguousarray(f["v_imu"], np.float64).tobytes())
        fh.write(np.array([f["pcl_beg_time"], f["pcl_end_time"], f["last_lidar_end_time"]], np.float64).tobytes())
        fh.write(np.concatenate([f["R"].ravel(), f["p"], f["vel"], f["bg"], f["ba"], f["grav"], f["cov"].ravel()]).astype(np.float64).tobytes())
        fh.write(np.concatenate([f["acc_s_last"], f["angvel_last"]]).astype(np.float64).tobytes())
        fh.write(np.concatenate([f["cov_gyr"], f["cov_acc"], f["cov_bias_gyr"], f["cov_bias_acc"], [f["G_m_s2"], f["mean_acc_norm"]],
                                 np.asarray(f["R_LI"]).ravel(), f["t_LI"]]).astype(np.float64).tobytes())
        fh.write(np.concatenate([f["pts"], f["offset_ms"][:, None]], 1).astype(np.float32).tobytes())
    r = subprocess.run([EXE_IMU, str(inp), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    st = np.frombuffer(raw, np.float64, 346, 0)
    pts = np.frombuffer(raw, np.float32, n * 3, 346 * 8).reshape(n, 3)
    P, C, x = oracle_inputs(po, f)
    ref_pts, _ = po.imu_undistort(P, C, f["v_imu"], f["pcl_beg_time"], f["pcl_end_time"], x, f["pts"], f["offset_ms"])
    ref = np.concatenate([x.rot[:], x.pos[:], x.vel[:], x.cov[:], [C.last_lidar_end_time], C.acc_s_last[:], C.angvel_last[:]])
    scale = np.maximum(np.abs(ref), 1e-12)
    assert (np.abs(st - ref) / np.maximum(scale, np.abs(ref[15:339]).max() * (np.arange(346) >= 15) * (np.arange(346) < 339))).max() < 1e-11
    assert np.abs(pts - ref_pts).max() <= 1e-5 and (pts == ref_pts).mean() > 0.999

"""pcl::VoxelGrid on the scan (first half; src/laserMapping.cpp:1398-1399 and the 0.2 m filter of
src/lidar_selection.cpp:351-352).  PCL is absent here, and its within-leaf summation order is unspecified, so the oracle is a
restatement with a DEFINED order (input order); CPU tier: properties that hold for every order; GPU tier: bit-exact vs it."""
import numpy as np
import pytest

from conftest import bits


def _cloud(flb, name="T1"):
    f = flb.synth.make_frame(name)
    return f["scan_body"]


@pytest.mark.parametrize("leaf", [0.2, 0.5])
def test_oracle_voxel_grid_properties(flb, po, leaf):
    xyz = _cloud(flb)
    out = po.voxel_grid(xyz, leaf)
    inv = np.float32(1.0) / np.float32(leaf)
    key = np.floor(xyz * inv).astype(np.int64)
    uniq, inverse, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    assert len(out) == len(uniq) and 1 < len(out) < len(xyz)
    # every output is the mean of exactly the points of one leaf (float64 mean within float32 rounding of the sum)
    okey = np.floor(out.astype(np.float64) * float(inv) + 1e-9 * np.sign(out)).astype(np.int64)
    mean = np.zeros((len(uniq), 3))
    np.add.at(mean, inverse.ravel(), xyz.astype(np.float64))
    mean /= counts[:, None]
    # PCL's order: ascending idx = i0 + i1 * div0 + i2 * div0 * div1 -> z-major, then y, then x
    order = np.lexsort((uniq[:, 0], uniq[:, 1], uniq[:, 2]))
    np.testing.assert_allclose(out, mean[order], rtol=0, atol=2e-5)
    # idempotence on the leaf structure: filtering the centroids again keeps one point per leaf
    again = po.voxel_grid(out, leaf)
    assert len(again) <= len(out)


@pytest.mark.gpu
@pytest.mark.parametrize("name,leaf", [("T1", 0.2), ("T1", 0.5), ("C2", 0.2)])
def test_gpu_voxel_grid_matches_oracle(flb, po, name, leaf):
    xyz = _cloud(flb, name)
    h = flb.Handle(device=0)
    got = h.voxel_grid(xyz, leaf)
    ref = po.voxel_grid(xyz, leaf)
    assert got.shape == ref.shape and (bits(got) == bits(ref)).all()
    # a strided input (PointXYZINormal: 12 floats per point) gives the same result
    wide = np.zeros((len(xyz), 12), np.float32)
    wide[:, :3] = xyz
    assert (bits(h.voxel_grid(wide, leaf)) == bits(ref)).all()
    with pytest.raises(flb.FlbError):
        h.voxel_grid(np.array([[0, 0, 0], [4e6, 4e6, 4e6]], np.float32), 0.001)      # PCL: leaf size too small
    h.close()

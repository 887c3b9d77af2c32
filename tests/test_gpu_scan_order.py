"""The upload's scan ordering: 24-bit Morton key (8 bits per axis, cell = twice the map grid's), STABLE in the caller's
index.  Two device paths produce it -- the one-block fused kernel (k_scan_sort_block, up to 25 600 points, three
instantiations) and keys + device-wide radix sort + gather above that -- and both must equal a numpy stable sort of
the same keys, bit for bit, at every size either side of an instantiation boundary.  Back-to-back uploads exercise the
two scan slots (the upload fills the slot that is not being read)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spread(v):
    v = v.astype(np.uint32) & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def _expected_order(xyz, cell_size):
    xyz = xyz.astype(np.float32)
    lo = xyz.min(axis=0)
    inv = np.float32(0.5) / np.float32(cell_size)
    c = np.clip((xyz - lo) * inv, np.float32(0), np.float32(255)).astype(np.uint32)
    key = _spread(c[:, 0]) | (_spread(c[:, 1]) << 1) | (_spread(c[:, 2]) << 2)
    return np.argsort(key, kind="stable").astype(np.int32)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 31, 512 * 17, 512 * 17 + 1, 512 * 33, 512 * 33 + 1, 24000, 1024 * 25, 1024 * 25 + 1, 40000])
def test_scan_order_matches_stable_morton(flb, n, mode):
    rng = np.random.default_rng(n)
    # few occupied cells => long runs of equal keys: the tie order (caller's index) is what is being tested
    xyz = (rng.integers(0, 6, size=(n, 3)) * 1.2 + rng.uniform(0, 0.3, size=(n, 3))).astype(np.float32)
    h = flb.Handle(cell_size=0.6)
    try:
        h.debug_set_scan_sort(mode)     # 0: by size / stream state; 1: the one-block kernel where it applies; 2: device-wide
        h.scan_upload(xyz)
        got = h.debug_scan_order()
        assert np.array_equal(got, _expected_order(xyz, 0.6))
        # a second and third upload land in the other slot / the first again
        xyz2 = xyz[::-1].copy()
        h.scan_upload(xyz2)
        assert np.array_equal(h.debug_scan_order(), _expected_order(xyz2, 0.6))
        h.scan_upload(xyz)
        assert np.array_equal(h.debug_scan_order(), got)
    finally:
        h.close()


def test_scan_order_wide_extent_clamps(flb):
    rng = np.random.default_rng(5)
    xyz = rng.uniform(-400, 400, size=(20000, 3)).astype(np.float32)      # > 255 cells per axis: keys clamp
    h = flb.Handle(cell_size=0.6)
    try:
        h.scan_upload(xyz)
        assert np.array_equal(h.debug_scan_order(), _expected_order(xyz, 0.6))
    finally:
        h.close()

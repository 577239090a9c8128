"""GPU parity: VoxelProcessor / VolumeCalculator kernels vs the CPU oracle (bit-exact integer work)."""
import numpy as np
import pytest
import torch
from scipy import ndimage

from conftest import random_blobs

pytestmark = pytest.mark.gpu


def dev_volume(eng, vol_bool):
    return eng.pack(eng.upload_u8(vol_bool.view(np.uint8)), 1)


@pytest.mark.parametrize("shape", [(3, 5, 32), (4, 7, 64), (2, 9, 1024), (5, 6, 31), (3, 4, 33), (2, 3, 100), (1, 1, 1),
                                   (7, 130, 257), (2, 2, 4096)])
@pytest.mark.parametrize("thr", [1, 200, 255])
def test_pack_unpack(eng, shape, thr):
    rng = np.random.default_rng(hash((shape, thr)) % 2**32)
    u8 = rng.integers(0, 256, size=shape, dtype=np.uint8)
    dv = eng.pack(eng.upload_u8(u8), thr)
    got = dv.to_host()
    assert got.dtype == np.bool_ and got.shape == shape
    assert np.array_equal(got, u8 >= thr)
    # bit layout: LSB-first along x, tail bits zero
    bits = dv.bits.cpu().numpy().view(np.uint32)
    W = shape[2]
    ref = np.zeros(shape[:2] + (eng.words_per_row(W) * 32,), dtype=np.uint8)
    ref[..., :W] = u8 >= thr
    refw = np.packbits(ref, axis=-1, bitorder="little").view(np.uint32)
    assert np.array_equal(bits, refw.reshape(bits.shape))


@pytest.mark.parametrize("shape", [(3, 1, 128), (5, 8, 128), (9, 3, 384), (70, 16, 256), (4, 33, 1024), (131, 5, 128)])
@pytest.mark.parametrize("thr", [1, 2, 127, 128, 129, 200, 254, 255])
def test_pack_gap_fused_kernel(eng, shape, thr):
    """t3d_pack_gap: threshold + stack + z gap fill (voxel_processor.py:46, 72-75) + per-slice counts + extrema in one pass;
    every byte value against every class of threshold (the byte compare is a hand-written carry trick, not a SIMD intrinsic)."""
    from tomography_3d_reconstructor_b200 import _lib
    lib = _lib.load()
    Z, H, W = shape
    rng = np.random.default_rng(hash((shape, thr)) % 2**32)
    u8 = rng.integers(0, 256, size=shape, dtype=np.uint8)
    u8[:, 0, :] = np.arange(W, dtype=np.int64).astype(np.uint8)[None, :] + np.arange(Z, dtype=np.uint8)[:, None]   # all byte values
    if Z > 4:
        u8[Z // 2] = 0                                  # a slice the gap fill has to bridge
    t = eng.upload_u8(u8)
    nw = eng.words_per_row(W)
    bits = torch.empty((Z, H, nw), dtype=torch.int32, device="cuda")
    cnt = torch.zeros((Z,), dtype=torch.int64, device="cuda")      # accumulated into: zeroed by the caller
    bb = torch.zeros((6,), dtype=torch.int32, device="cuda")
    rc = lib.t3d_pack_gap(eng._p(t), Z, H, W, thr, eng._p(bits), eng._p(cnt), eng._p(bb), eng._stream())
    assert rc == 0, lib.t3d_last_error()
    torch.cuda.synchronize()
    v = u8 >= thr
    ref = v.copy()
    ref[1:-1] |= v[:-2] & v[2:]
    got = np.unpackbits(bits.cpu().numpy().view(np.uint8).reshape(Z, H, nw * 4), axis=-1, bitorder="little")[..., :W].astype(bool)
    assert np.array_equal(got, ref)
    assert np.array_equal(cnt.cpu().numpy(), ref.reshape(Z, -1).sum(axis=1))
    b = bb.cpu().numpy().view(np.uint32).astype(np.int64)
    if ref.any():
        zz, yy, xx = np.nonzero(ref)
        want = [zz.min(), zz.max(), yy.min(), yy.max(), xx.min(), xx.max()]
        have = [0x7fffffff - b[0], b[1] - 1, 0x7fffffff - b[2], b[3] - 1, 0x7fffffff - b[4], b[5] - 1]
        assert have == [int(x) for x in want]
    # unsupported stacks are refused, not mangled
    assert lib.t3d_pack_gap(eng._p(t), 2, H, W, thr, eng._p(bits), None, None, eng._stream()) == 2


def test_pack_unaligned_input(eng):
    rng = np.random.default_rng(5)
    big = rng.integers(0, 2, size=(4 * 8 * 64 + 3,), dtype=np.uint8)
    t = torch.from_numpy(big).cuda()[3:].view(4, 8, 64)  # 3-byte offset: generic path
    dv = eng.pack(t, 1)
    assert np.array_equal(dv.to_host(), big[3:].reshape(4, 8, 64) >= 1)


@pytest.mark.parametrize("shape", [(40, 48), (64, 64), (33, 100), (128, 1024), (5, 5), (1, 7), (200, 70), (70, 2100)])
def test_fill_holes_random(eng, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for density in (0.3, 0.5, 0.7):
        planes = np.stack([random_blobs(rng, shape, density, 2.0) for _ in range(2)])
        vol = np.concatenate([planes[:1], np.zeros((1,) + shape, bool), planes[1:]])
        dv = dev_volume(eng, vol)
        out = eng.close_volume_ends(dv)  # z=1 slice gets a&b of the filled ends
        f0, f2 = ndimage.binary_fill_holes(vol[0]), ndimage.binary_fill_holes(vol[2])
        got = out.to_host()
        assert np.array_equal(got[0], f0)
        assert np.array_equal(got[2], f2)
        assert np.array_equal(got[1], f0 & f2)


def test_fill_holes_known_answer_and_hard_shapes(eng):
    # scipy docstring example (scipy/ndimage/_morphology.py binary_fill_holes)
    a = np.zeros((5, 5), dtype=bool)
    a[1:4, 1:4] = True
    a[2, 2] = False
    cases = [a]
    # spiral wall: background is a long snake, converges only after many row/column closures
    n = 65
    s = np.zeros((n, n), bool)
    lo, hi = 1, n - 2
    while hi - lo > 3:
        s[lo, lo:hi + 1] = True
        s[lo:hi + 1, hi] = True
        s[hi, lo + 2:hi + 1] = True
        s[lo + 2:hi + 1, lo + 2] = True
        lo += 4
        hi -= 4
    cases.append(s)
    cases.append(np.ones((9, 40), bool))
    cases.append(np.zeros((9, 40), bool))
    ring = np.zeros((50, 300), bool)
    ring[5:45, 10:290] = True
    ring[10:40, 20:280] = False
    ring[20:30, 100:200] = True
    ring[23:27, 120:180] = False
    cases.append(ring)
    for c in cases:
        vol = np.stack([c, c])
        got = eng.close_volume_ends(dev_volume(eng, vol)).to_host()
        assert np.array_equal(got[0], ndimage.binary_fill_holes(c))
        assert np.array_equal(got[1], ndimage.binary_fill_holes(c))


@pytest.mark.parametrize("shape", [(1, 8, 40), (2, 8, 40), (3, 9, 33), (8, 20, 70), (17, 33, 129)])
def test_close_volume_ends_vs_oracle(eng, oracle, shape):
    rng = np.random.default_rng(shape[0])
    for density in (0.2, 0.5, 0.8):
        vol = rng.random(shape) < density
        got = eng.close_volume_ends(dev_volume(eng, vol))
        ref = oracle.close_volume_ends(vol)
        assert np.array_equal(got.to_host(), ref)
        assert np.array_equal(got.slice_counts(), ref.reshape(shape[0], -1).sum(axis=1))


STAGE_SETS = [[True], [False], [True, False], [False, True], [True, False, False, True], [False, True, True, False],
              [True, True, True, True], [False, False, False], [True, False, False, True, False, True]]


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 5), (5, 7, 31), (9, 17, 33), (8, 16, 256), (11, 35, 300), (20, 40, 64)])
def test_morph_vs_scipy(eng, oracle, shape):
    rng = np.random.default_rng(shape[1] * 7 + shape[2])
    vol = random_blobs(rng, shape, 0.5, 1.0) if min(shape) > 2 else rng.random(shape) < 0.5
    dv = dev_volume(eng, vol)
    for stages in STAGE_SETS:
        ref = vol
        for er in stages:
            ref = oracle.binary_erosion6(ref) if er else oracle.binary_dilation6(ref)
        got = eng.morph(dv, stages)
        assert np.array_equal(got.to_host(), ref), stages
        assert np.array_equal(got.slice_counts(), ref.reshape(shape[0], -1).sum(axis=1)), stages


@pytest.mark.parametrize("iterations,manifold", [(3, True), (1, True), (0, True), (3, False), (0, False)])
def test_smooth_vs_oracle(eng, oracle, iterations, manifold):
    rng = np.random.default_rng(11)
    vol = random_blobs(rng, (14, 40, 72), 0.45, 1.2)
    got = eng.smooth(dev_volume(eng, vol), iterations, manifold).to_host()
    assert np.array_equal(got, oracle.smooth_voxel_data(vol, iterations, manifold))


def test_stats_and_bbox(eng):
    rng = np.random.default_rng(3)
    vol = np.zeros((9, 50, 130), bool)
    vol[2:7, 10:33, 40:101] = rng.random((5, 23, 61)) < 0.3
    vol[2, 10, 40] = vol[6, 32, 100] = True
    dv = dev_volume(eng, vol)
    assert np.array_equal(dv.slice_counts(), vol.reshape(9, -1).sum(axis=1))
    z, y, x = np.where(vol)
    assert dv.bbox() == (z.min(), z.max(), y.min(), y.max(), x.min(), x.max())
    assert dev_volume(eng, np.zeros((3, 4, 5), bool)).bbox() is None


@pytest.mark.parametrize("sub", [1, 2, 5])
def test_point_cloud(eng, oracle, sub):
    rng = np.random.default_rng(4)
    vol = rng.random((6, 20, 70)) < 0.2
    depths = oracle.calculate_slice_depths(6.0, 1, 4, 1)
    got = eng.point_cloud(dev_volume(eng, vol), 0.28, 0.19, depths, sub)
    ref = oracle.generate_point_cloud(vol, 0.28, 0.19, depths, sub)
    assert got.dtype == np.float64 and got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_classes_vs_oracle(eng, oracle):
    from tomography_3d_reconstructor_b200 import VoxelProcessor, VolumeCalculator
    Z, H, W = 20, 48, 96
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, 20:24, 40:50] = 0
    masks = [u8[z] >= 200 for z in range(Z)]
    vp, vc = VoxelProcessor(), VolumeCalculator()
    with pytest.raises(ValueError, match="Load masks first"):
        vp.create_voxel_data([])
    raw = vp.create_voxel_data(masks, False)
    assert np.array_equal(raw, np.stack(masks))
    vox = vp.create_voxel_data(masks, True, 4, 12, 4)
    ref = oracle.create_voxel_data(masks, True)
    assert vox.dtype == np.bool_ and np.array_equal(vox, ref) and vp.voxel_data is vox
    assert (vp.side_0_count, vp.side_1_count, vp.side_2_count) == (4, 12, 4)
    depths = vp.calculate_slice_depths(6.0)
    assert np.array_equal(depths, oracle.calculate_slice_depths(6.0, 4, 12, 4))
    mm_x, mm_y = 143.1 / W, 95.03 / H
    assert vc.calculate_voxel_volume_variable_depth(vox, mm_x, mm_y, depths) == \
        oracle.calculate_voxel_volume_variable_depth(ref, mm_x, mm_y, depths)
    assert vc.calculate_voxel_volume(vox, mm_x, mm_y, 0.1) == oracle.calculate_voxel_volume(ref, mm_x, mm_y, 0.1)
    assert vc.calculate_bounding_box_variable_depth(vox, mm_x, mm_y, depths) == \
        oracle.calculate_bounding_box_variable_depth(ref, mm_x, mm_y, depths)
    assert vc.calculate_bounding_box(vox, mm_x, mm_y, 0.1) == oracle.calculate_bounding_box(ref, mm_x, mm_y, 0.1)
    # arrays that did not come from this package are uploaded and give the same answers
    fresh = ref.copy()
    assert vc.calculate_voxel_volume_variable_depth(fresh, mm_x, mm_y, depths) == \
        oracle.calculate_voxel_volume_variable_depth(ref, mm_x, mm_y, depths)
    # occupancies in another dtype (0 / 1 values): accepted like the reference's np.sum / np.where accept them
    for dt in (np.uint8, np.int32, np.float32):
        other = ref.astype(dt)
        assert vc.calculate_voxel_volume(other, mm_x, mm_y, 0.1) == oracle.calculate_voxel_volume(ref, mm_x, mm_y, 0.1)
        assert vc.calculate_bounding_box(other, mm_x, mm_y, 0.1) == oracle.calculate_bounding_box(ref, mm_x, mm_y, 0.1)
    with pytest.raises(TypeError, match="only 0 and 1"):      # 0 / 255 masks are not occupancies: refused, not reinterpreted
        vc.calculate_voxel_volume(ref.astype(np.uint8) * 255, mm_x, mm_y, 0.1)
    sm = vp.smooth_voxel_data(fresh, 3, True)
    assert np.array_equal(sm, oracle.smooth_voxel_data(ref, 3, True))
    pc = vp.generate_point_cloud(sm, mm_x, mm_y, depths, 2)
    assert np.array_equal(pc, oracle.generate_point_cloud(np.asarray(sm), mm_x, mm_y, depths, 2))

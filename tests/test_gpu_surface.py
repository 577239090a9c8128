"""GPU parity: SurfaceExtractor kernels vs the CPU oracle.

Bars (BASELINE.json north_star): field sign / cube cases / triangle counts bit-exact; vertex positions within
1e-5 relative (1e-3 voxel absolute); mesh volume within 1e-6 relative after a canonical sort of faces.
"""
import ctypes

import numpy as np
import pytest
import torch
from scipy import ndimage

from conftest import random_blobs

pytestmark = pytest.mark.gpu


def dev_volume(eng, vol_bool):
    return eng.pack(eng.upload_u8(np.ascontiguousarray(vol_bool).view(np.uint8)), 1)


def field_dense(eng, dv, pad, gaussian):
    Z, H, W = dv.shape
    out = torch.empty((Z + 2 * pad, H + 2 * pad, W + 2 * pad), dtype=torch.float32, device="cuda")
    eng.check(eng._L().t3d_field_dense(eng._p(dv.bits), Z, H, W, pad, gaussian, eng._W3_C, eng._p(out), eng._stream()),
              "t3d_field_dense")
    return out.cpu().numpy()


def sign_dense(sign, dims):
    Zs, Hs, Ws = dims
    bits = np.unpackbits(sign.cpu().numpy().view(np.uint8), axis=-1, bitorder="little")
    return bits.reshape(Zs, Hs, -1)[:, :, :Ws].astype(bool)


def test_gaussian_weights_match_scipy(eng):
    from scipy.ndimage._filters import _gaussian_kernel1d
    k = _gaussian_kernel1d(0.5, 0, 2)
    assert np.array_equal(eng._W3, np.array([k[2], k[1], k[0]]))
    assert eng._W3[0].hex() == "0x1.92b965ef5aaeep-1" and eng._W3[1].hex() == "0x1.b405b9842b206p-4"
    assert eng._W3[2].hex() == "0x1.14aebe6a24088p-12"


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 2, 3), (2, 2, 2), (3, 4, 5), (6, 9, 31), (7, 12, 33), (9, 20, 70), (12, 33, 130)])
@pytest.mark.parametrize("pad", [0, 1])
def test_field_bit_exact_vs_scipy(eng, oracle, shape, pad):
    rng = np.random.default_rng(shape[2] * 3 + pad)
    for density in (0.15, 0.5, 0.85):
        vol = rng.random(shape) < density
        dv = dev_volume(eng, vol)
        got = field_dense(eng, dv, pad, 1)
        ref = oracle.scalar_field(vol, True, bool(pad))
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "float32 field differs bitwise from scipy"
        sign, dims, n_exact = eng.field_sign(dv, pad)
        assert np.array_equal(sign_dense(sign, dims), ref > 0.5)


def test_field_sign_needs_exact_path(eng, oracle):
    # isolated voxels / isolated holes are the only places where sign != occupancy
    vol = np.zeros((9, 9, 40), bool)
    vol[4, 4, 4] = True                      # isolated voxel: 0.4866 -> outside
    vol[4, 4, 10] = vol[3, 3, 11] = vol[5, 5, 11] = vol[3, 5, 9] = vol[5, 3, 9] = True  # diagonal cluster
    vol[2:7, 2:7, 20:30] = True
    vol[4, 4, 25] = False                    # isolated hole
    dv = dev_volume(eng, vol)
    sign, dims, n_exact = eng.field_sign(dv, 1)
    ref = oracle.scalar_field(vol, True, True)
    assert int(n_exact.item()) >= 2
    assert np.array_equal(sign_dense(sign, dims), ref > 0.5)
    assert not np.array_equal(sign_dense(sign, dims)[1:-1, 1:-1, 1:-1], vol)


@pytest.mark.parametrize("shape,pad", [((6, 9, 31), 1), ((7, 12, 33), 0), ((10, 30, 70), 1), ((5, 40, 129), 1)])
def test_cube_cases_bit_exact(eng, oracle, shape, pad):
    rng = np.random.default_rng(shape[1])
    vol = random_blobs(rng, shape, 0.5, 1.0)
    dv = dev_volume(eng, vol)
    sign, (Zs, Hs, Ws), _ = eng.field_sign(dv, pad)
    out = torch.empty((Zs - 1, Hs - 1, Ws - 1), dtype=torch.uint8, device="cuda")
    eng.check(eng._L().t3d_cube_cases(eng._p(sign), Zs, Hs, Ws, eng._p(out), eng._stream()), "t3d_cube_cases")
    ref = oracle.cube_cases(oracle.scalar_field(vol, True, bool(pad)), 0.5)
    assert np.array_equal(out.cpu().numpy(), ref)


def canon_faces(f):
    """rotate each face so its smallest index comes first (keeps winding), then sort rows"""
    f = np.asarray(f, dtype=np.int64)
    k = np.argmin(f, axis=1)
    r = np.stack([np.take_along_axis(f, ((k + i) % 3)[:, None], 1)[:, 0] for i in range(3)], axis=1)
    return r[np.lexsort((r[:, 2], r[:, 1], r[:, 0]))]


def check_mesh(got, ref, n_amb_ok=True):
    v, f = got
    rv, rf = ref
    assert v.dtype == np.float32 and v.shape == rv.shape
    assert f.shape == rf.shape
    assert np.allclose(v, rv, rtol=1e-5, atol=1e-5)
    assert np.array_equal(f, rf)  # same order as the reference, not just the same set


PHANTOMS = [((24, 64, 96), (3, 18, 3)), ((40, 50, 70), (5, 30, 5)), ((16, 33, 130), (0, 16, 0))]


@pytest.mark.parametrize("shape,sides", PHANTOMS)
@pytest.mark.parametrize("add_padding", [True, False])
def test_extract_phantom_vs_oracle(eng, oracle, shape, sides, add_padding):
    from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor
    Z, H, W = shape
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    vol = oracle.smooth_voxel_data(oracle.create_voxel_data([u8[z] >= 200 for z in range(Z)]), 3, True)
    depths = oracle.calculate_slice_depths(6.0, *sides)
    mm_x, mm_y = 143.1 / W, 95.03 / H
    se = SurfaceExtractor()
    got = se.extract_manifold_surface(vol, depths, mm_y, mm_x, True, True, add_padding)
    assert got is not None, se.last_error
    rv, rf, namb = oracle.extract_manifold_surface(vol, depths, mm_y, mm_x, True, True, add_padding, return_diag=True)
    assert namb == 0 and se.last_n_ambiguous == 0
    check_mesh(got, (rv, rf))
    assert got[1].dtype == np.int64
    # strict lexicographic order of the unique vertices (np.unique contract)
    v = got[0]
    d = np.diff(v.astype(np.float64), axis=0)
    lex = (d[:, 0] > 0) | ((d[:, 0] == 0) & ((d[:, 1] > 0) | ((d[:, 1] == 0) & (d[:, 2] > 0))))
    assert lex.all()
    mv = se.calculate_mesh_volume(*got)
    ref_mv = oracle.calculate_mesh_volume_f64(rv, rf)
    assert abs(mv - ref_mv) <= 1e-6 * ref_mv
    ar = se.calculate_surface_area(*got)
    ref_ar = oracle.calculate_surface_area_f64(rv, rf)
    assert abs(ar - ref_ar) <= 1e-6 * ref_ar
    # fresh host arrays (not from the registry) take the upload path and agree
    mv2 = se.calculate_mesh_volume(got[0].copy(), got[1].copy())
    assert mv2 == mv


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("add_padding", [True, False])
def test_extract_noisy_vs_oracle(eng, oracle, seed, add_padding):
    """General inputs (incl. ambiguous cubes, objects touching the border, isolated voxels)."""
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    rng = np.random.default_rng(seed)
    vol = random_blobs(rng, (13, 37, 75), 0.4, 0.8)
    vol |= rng.random(vol.shape) < 0.01
    depths = oracle.calculate_slice_depths(6.0, 2, 9, 2)
    se = SurfaceExtractor()
    got = se.extract_manifold_surface(vol, depths, 0.31, 0.27, True, True, add_padding)
    assert got is not None, se.last_error
    rv, rf, namb = oracle.extract_manifold_surface(vol, depths, 0.31, 0.27, True, True, add_padding, return_diag=True)
    assert se.last_n_ambiguous == namb
    check_mesh(got, (rv, rf))


def test_extract_non_manifold_mode(eng, oracle):
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    rng = np.random.default_rng(9)
    vol = random_blobs(rng, (10, 30, 66), 0.5, 1.5)
    depths = oracle.calculate_slice_depths(6.0, 2, 6, 2)
    se = SurfaceExtractor()
    got = se.extract_manifold_surface(vol, depths, 0.2, 0.3, True, False, True)
    assert got is not None, se.last_error
    rv, rf = oracle.extract_manifold_surface(vol, depths, 0.2, 0.3, True, False, True)
    # no np.unique here: vertex numbering is emission order, which differs from skimage's first-use order;
    # compare geometry per face corner instead
    gv, gf = got
    assert gv.shape == rv.shape and gf.shape == rf.shape and gf.dtype == np.int32
    assert np.allclose(gv[gf], rv[rf], rtol=1e-5, atol=1e-5)


def test_extract_failure_modes(eng, oracle):
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    se = SurfaceExtractor()
    depths = np.full(4, 0.5)
    assert se.extract_manifold_surface(np.zeros((4, 8, 8), bool), depths, 1.0, 1.0) is None          # empty volume
    assert oracle.extract_manifold_surface(np.zeros((4, 8, 8), bool), depths, 1.0, 1.0) is None
    assert se.extract_manifold_surface(np.ones((4, 8, 8), bool), depths, 1.0, 1.0, True, True, False) is None  # no crossing
    assert oracle.extract_manifold_surface(np.ones((4, 8, 8), bool), depths, 1.0, 1.0, True, True, False) is None
    one = np.zeros((4, 8, 8), bool)
    one[1, 3, 3] = True  # an isolated voxel blurs to 0.4866 < 0.5: no surface
    assert se.extract_manifold_surface(one, depths, 1.0, 1.0) is None
    assert oracle.extract_manifold_surface(one, depths, 1.0, 1.0) is None
    got = se.extract_manifold_surface(np.ones((4, 8, 8), bool), depths, 1.0, 1.0)  # full volume, padded: a box
    ref = oracle.extract_manifold_surface(np.ones((4, 8, 8), bool), depths, 1.0, 1.0)
    check_mesh(got, ref)
    # empty slice_depths: z stays in index units (surface_extractor.py:84-86)
    got = se.extract_manifold_surface(np.ones((4, 8, 8), bool), np.array([]), 1.0, 1.0)
    ref = oracle.extract_manifold_surface(np.ones((4, 8, 8), bool), np.array([]), 1.0, 1.0)
    check_mesh(got, ref)


def test_canonicalize_vs_numpy_unique(eng, oracle):
    rng = np.random.default_rng(7)
    V, F = 5000, 9000
    base = rng.integers(-3, 4, size=(400, 3)).astype(np.float32) * np.float32(0.37)
    base[5] = [0.0, -0.0, 0.0]
    verts = base[rng.integers(0, 400, V)]
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int32)
    gv, gf = eng.canonicalize(torch.from_numpy(verts).cuda(), torch.from_numpy(faces).cuda())
    rv, rf = oracle.ensure_manifold_mesh(verts, faces)
    assert np.array_equal(gv.cpu().numpy(), rv)
    assert np.array_equal(gf.cpu().numpy(), rf)


def test_mesh_measures_vs_f64(eng, oracle):
    rng = np.random.default_rng(8)
    verts = rng.random((3000, 3)).astype(np.float32) * 50
    faces = rng.integers(0, 3000, size=(20000, 3)).astype(np.int64)
    vol, area = eng.mesh_measure(torch.from_numpy(verts).cuda(), torch.from_numpy(faces).cuda())
    v = verts.astype(np.float64)
    a, b, c = v[faces[:, 0]], v[faces[:, 1]], v[faces[:, 2]]
    ref_vol = np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0
    assert abs(vol - ref_vol) <= 1e-9 * np.abs(np.einsum("ij,ij->i", a, np.cross(b, c))).sum()
    assert abs(area - oracle.calculate_surface_area_f64(verts, faces)) <= 1e-9 * area


def test_watertight_at_moderate_size(eng, oracle):
    """Size-independent properties on a volume the oracle would take long to process."""
    from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor, VolumeCalculator
    Z, H, W = 96, 256, 320
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    vp, se, vc = VoxelProcessor(), SurfaceExtractor(), VolumeCalculator()
    vox = vp.create_voxel_data([u8[z] >= 200 for z in range(Z)], True, 12, 72, 12)
    sm = vp.smooth_voxel_data(vox, 3, True)
    depths = vp.calculate_slice_depths(6.0)
    v, f = se.extract_manifold_surface(sm, depths, 95.03 / H, 143.1 / W)
    assert se.last_n_ambiguous == 0
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    key = e[:, 0] * (len(v) + 1) + e[:, 1]
    rev = e[:, 1] * (len(v) + 1) + e[:, 0]
    assert len(np.unique(key)) == len(key)                     # every directed edge once
    assert np.array_equal(np.sort(key), np.sort(rev))          # and its twin exists: closed, oriented
    assert len(v) - len(key) // 2 + len(f) == 2                # Euler characteristic of a sphere
    # uniform-depth analytic check: mesh volume ~ voxel volume ~ 4/3 pi abc (index units)
    ones = np.ones(Z)
    v1, f1 = se.extract_manifold_surface(sm, ones, 1.0, 1.0, True, True, False)
    mv = se.calculate_mesh_volume(v1, f1)
    analytic = 4.0 / 3.0 * np.pi * (0.42 * Z) * (0.33 * H) * (0.45 * W)
    assert abs(mv - analytic) / analytic < 5e-3
    assert abs(vc.calculate_voxel_volume(sm, 1.0, 1.0, 1.0) - analytic) / analytic < 5e-3


def test_mm_scaling_follows_numpy_promotion(eng, oracle):
    """`vertices[:, 1] *= mm`: float32 multiply for a python float, float64 multiply for a numpy float64 scalar."""
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    vol = oracle.smooth_voxel_data(oracle.ellipsoid_phantom_u8(20, 40, 60) >= 200, 3, True)
    depths = oracle.calculate_slice_depths(6.0, 2, 16, 2)
    se = SurfaceExtractor()
    for mm_y, mm_x in ((95.03 / 40, 143.1 / 60), (np.float64(95.03 / 40), np.float64(143.1 / 60))):
        got = se.extract_manifold_surface(vol, depths, mm_y, mm_x)
        ref = oracle.extract_manifold_surface(vol, depths, mm_y, mm_x)
        assert np.array_equal(got[0], ref[0]) or np.allclose(got[0], ref[0], rtol=1e-6, atol=0)
        assert np.array_equal(got[0][:, 0], ref[0][:, 0])       # z map is bit-exact
        assert np.array_equal(got[1], ref[1])


def test_fast_canonicalize_falls_back_when_order_cannot_be_verified(eng, oracle):
    rng = np.random.default_rng(17)
    V, F = 4000, 7000
    base = rng.integers(-3, 4, size=(300, 3)).astype(np.float32) * np.float32(0.41)
    verts = base[rng.integers(0, 300, V)]                      # arbitrary order: the (z,y)-key shortcut is invalid here
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int32)
    tv, tf = torch.from_numpy(verts).cuda(), torch.from_numpy(faces).cuda()
    _, _, counts = eng.canonicalize(tv, tf, sync=False, fast=True)
    assert int(counts[2].item()) != 0                          # detected on the device ...
    gv, gf = eng.canonicalize(tv, tf, fast=True)               # ... and the synchronous call falls back
    rv, rf = oracle.ensure_manifold_mesh(verts, faces)
    assert np.array_equal(gv.cpu().numpy(), rv) and np.array_equal(gf.cpu().numpy(), rf)


@pytest.mark.parametrize("dup_every", [0, 7, 1])
def test_unique_pass_with_and_without_coinciding_vertices(eng, oracle, dup_every):
    """The np.unique step runs an optimistic streaming pass (no two vertices coincide => index = sorted position) and falls
    back on the device to the exact scan kernel when it meets an equal pair: both against numpy on vertex lists that are in
    sorted order already (so the fast ordering is verified and the result of the unique kernels is what comes back)."""
    rng = np.random.default_rng(23 + dup_every)
    pts = np.unique(rng.integers(-40, 41, size=(6000, 3)).astype(np.float32) * np.float32(0.173), axis=0)   # sorted by (z, y, x)
    if dup_every:
        reps = np.where(np.arange(len(pts)) % dup_every == 0, 1 + (np.arange(len(pts)) % 3), 1)
        pts = np.repeat(pts, reps, axis=0)                   # runs of 1..3 equal vertices, still sorted
    V, F = len(pts), 2 * len(pts)
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int32)
    tv, tf = torch.from_numpy(pts).cuda(), torch.from_numpy(faces).cuda()
    vout, fout, counts = eng.canonicalize(tv, tf, sync=False, fast=True)
    v2, f2, bad = (int(c) for c in counts.cpu().tolist())
    assert bad == 0                                          # order verified on the device: no generic fallback involved
    rv, rf = oracle.ensure_manifold_mesh(pts, faces)
    assert v2 == len(rv) and f2 == len(rf)
    assert np.array_equal(vout[:v2].cpu().numpy(), rv) and np.array_equal(fout[:f2].cpu().numpy(), rf)


@pytest.mark.parametrize("seed,density", [(5, 0.35), (6, 0.5), (7, 0.65)])
def test_ambiguous_cubes_are_resolved_like_the_oracle(eng, oracle, seed, density):
    """Raw noise (no smoothing): thousands of cubes with ambiguous faces / case 4.  Lewiner's face and interior tests pick
    the tiling; GPU and oracle take the same decisions (same corner values, same arithmetic) -> identical faces."""
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    rng = np.random.default_rng(seed)
    occ = rng.random((12, 30, 70)) < density
    depths = oracle.calculate_slice_depths(6.0, 2, 8, 2)
    se = SurfaceExtractor()
    for add_padding in (True, False):
        got = se.extract_manifold_surface(occ, depths, 0.31, 0.27, True, True, add_padding)
        assert got is not None, se.last_error
        rv, rf, namb = oracle.extract_manifold_surface(occ, depths, 0.31, 0.27, True, True, add_padding, return_diag=True)
        amb, changed, tunnels, interior = oracle.last_mc33_stats
        assert namb > 1000 and changed > 100 and tunnels > 10, (namb, changed, tunnels)
        assert se.last_n_ambiguous == namb
        check_mesh(got, (rv, rf))


def test_ambiguous_cubes_on_a_dense_float_field(eng, oracle):
    """The dense-field entry (SDF path) resolves ambiguous cubes from the field values themselves."""
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    rng = np.random.default_rng(3)
    field = np.zeros((11, 21, 40), np.float32)
    field[1:-1, 1:-1, 1:-1] = rng.random((9, 19, 38)).astype(np.float32)
    depths = np.full(11, 0.5)
    se = SurfaceExtractor()
    got = se.extract_surface_from_sdf(field, depths, 0.3, 0.25, level=0.5)
    assert got is not None, se.last_error
    rv, rf, namb = oracle.marching_cubes(field, 0.5)
    amb, changed, tunnels, interior = oracle.last_mc33_stats
    assert changed > 100 and tunnels > 10
    oracle.apply_variable_slice_depths(rv, depths, False)
    rv[:, 1] *= 0.3
    rv[:, 2] *= 0.25
    uv, uf = oracle.ensure_manifold_mesh(rv, rf)
    assert se.last_n_ambiguous == namb
    assert np.array_equal(got[0], uv) and np.array_equal(got[1], uf)

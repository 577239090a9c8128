"""GPU parity of the additive exact EDT / SDF stage against scipy.ndimage.distance_transform_edt."""
import numpy as np
import pytest
import torch
from scipy import ndimage

from conftest import random_blobs

pytestmark = pytest.mark.gpu


def dev_volume(eng, vol_bool):
    return eng.pack(eng.upload_u8(np.ascontiguousarray(vol_bool).view(np.uint8)), 1)


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 4, 5), (9, 17, 33), (12, 40, 70), (20, 33, 130), (6, 5, 1100)])
@pytest.mark.parametrize("sampling", [(1.0, 1.0, 1.0), (0.09375, 0.31, 0.28)])
def test_edt_matches_scipy(eng, shape, sampling):
    from tomography_3d_reconstructor_b200 import edt
    rng = np.random.default_rng(shape[2])
    for density in (0.3, 0.9):
        vol = random_blobs(rng, shape, density, 1.5) if min(shape) > 2 else rng.random(shape) < density
        if vol.all():
            vol.flat[0] = False
        got = edt.distance(dev_volume(eng, vol), sampling).cpu().numpy()
        ref = ndimage.distance_transform_edt(vol, sampling=sampling)
        if sampling == (1.0, 1.0, 1.0):
            assert np.array_equal(got, ref.astype(np.float32))                 # exact integer squared distances
        else:
            assert np.allclose(got, ref, rtol=1e-6, atol=0)


def test_sdf_and_degenerate_volumes(eng, oracle):
    from tomography_3d_reconstructor_b200 import edt, VoxelProcessor
    vol = oracle.smooth_voxel_data(oracle.ellipsoid_phantom_u8(24, 48, 80) >= 200, 3, True)
    got = edt.signed_distance(dev_volume(eng, vol)).cpu().numpy()
    assert np.array_equal(got, oracle.signed_distance(vol))
    assert (got[vol] > 0).all() and (got[~vol] < 0).all()
    ones = np.ones((3, 4, 5), bool)
    assert np.isposinf(edt.signed_distance(dev_volume(eng, ones)).cpu().numpy()).all()
    assert np.isneginf(edt.signed_distance(dev_volume(eng, ~ones)).cpu().numpy()).all()
    vp = VoxelProcessor()
    assert np.array_equal(vp.compute_sdf(vol, (0.5, 1.0, 2.0)), oracle.signed_distance(vol, (0.5, 1.0, 2.0)))


def test_vertex_normals(eng, oracle):
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    vol = oracle.smooth_voxel_data(oracle.ellipsoid_phantom_u8(24, 48, 64) >= 200, 3, True)
    se = SurfaceExtractor()
    v, f = se.extract_manifold_surface(vol, np.ones(24), 1.0, 1.0)
    n = se.vertex_normals(v, f)
    assert n.shape == v.shape and np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    vv = v.astype(np.float64)
    fn = np.cross(vv[f[:, 1]] - vv[f[:, 0]], vv[f[:, 2]] - vv[f[:, 0]])
    acc = np.zeros_like(vv)
    for k in range(3):
        np.add.at(acc, f[:, k], fn)
    acc /= np.linalg.norm(acc, axis=1, keepdims=True)
    assert np.allclose(n, acc, atol=1e-4)
    centre = vv.mean(axis=0)                     # one consistent orientation relative to the centroid
    s = np.sign(np.einsum("ij,ij->i", n, vv - centre))
    assert abs(s.mean()) > 0.99


def test_surface_from_sdf(eng, oracle):
    """Additive path: exact SDF -> marching cubes at level 0 on the float field itself, against the oracle's MC."""
    from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor
    vol = oracle.smooth_voxel_data(oracle.ellipsoid_phantom_u8(24, 56, 72) >= 200, 3, True)
    depths = oracle.calculate_slice_depths(6.0, 3, 18, 3)
    vp, se = VoxelProcessor(), SurfaceExtractor()
    sdf = vp.compute_sdf(vol)
    assert np.array_equal(sdf, oracle.signed_distance(vol))
    got = se.extract_surface_from_sdf(sdf, depths, 0.3, 0.25, level=0.0)
    assert got is not None, se.last_error
    rv, rf, namb = oracle.marching_cubes(sdf, 0.0)
    oracle.apply_variable_slice_depths(rv, depths, False)
    rv[:, 1] *= 0.3
    rv[:, 2] *= 0.25
    uv, uf = oracle.ensure_manifold_mesh(rv, rf)
    assert se.last_n_ambiguous == namb
    assert np.array_equal(got[0], uv) and np.array_equal(got[1], uf)
    e = np.concatenate([uf[:, [0, 1]], uf[:, [1, 2]], uf[:, [2, 0]]])
    if namb == 0:
        key = e[:, 0] * (len(uv) + 1) + e[:, 1]
        assert np.array_equal(np.sort(key), np.sort(e[:, 1] * (len(uv) + 1) + e[:, 0]))   # closed, oriented


def _emulated_all_to_all(sends, send_sizes, recvs, recv_sizes):
    """What edt._exchange does over NCCL, for all ranks living in this process: chunk q of rank r -> chunk r of rank q."""
    world = len(sends)
    for r in range(world):
        so = np.concatenate([[0], np.cumsum(send_sizes[r])])
        for q in range(world):
            ro = np.concatenate([[0], np.cumsum(recv_sizes[q])])
            assert send_sizes[r][q] == recv_sizes[q][r]
            recvs[q][ro[r]:ro[r + 1]] = sends[r][so[q]:so[q + 1]]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("shape,sampling", [((19, 23, 40), (1.0, 1.0, 1.0)), ((24, 17, 70), (0.7, 1.3, 0.45))])
def test_sharded_edt_matches_single_device(eng, shape, sampling, world):
    """z-slab sharded signed distance (x/y passes per slab, all-to-all transpose, z pass per y-slab, transpose back), all
    ranks emulated on ONE GPU, must equal the single-device transform bit for bit."""
    from tomography_3d_reconstructor_b200 import edt, sharded
    rng = np.random.default_rng(5)
    Z, H, W = shape
    vol = random_blobs(rng, shape, 0.45, 1.5)
    dv = dev_volume(eng, vol)
    ref = edt.signed_distance(dv, sampling).cpu().numpy()
    ranges = [sharded.slab_range(Z, r, world) for r in range(world)]
    ts = [edt.SlabTransform(dv.bits[a:b].contiguous(), Z, a, H, W, sampling, r, world) for r, (a, b) in enumerate(ranges)]
    sends = [t.sdf_xy_pass() for t in ts]             # both polarities in one sweep: one offset pair per voxel is exchanged
    for c in range(2):
        _emulated_all_to_all([s[c] for s in sends], [t.send_sizes for t in ts], [t.cols[c] for t in ts], [t.recv_sizes for t in ts])
    for t in ts:
        t.sdf_z_pass()
    _emulated_all_to_all([t.dist_cols for t in ts], [t.recv_sizes for t in ts], [t.back for t in ts], [t.send_sizes for t in ts])
    out = torch.cat([t.result() for t in ts]).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    # the one-sided transform (edt(occ) only) still goes through the two-component exchange of t3d_edt_xy / t3d_edt_z
    ts = [edt.SlabTransform(dv.bits[a:b].contiguous(), Z, a, H, W, sampling, r, world) for r, (a, b) in enumerate(ranges)]
    sends = [t.xy_pass(0) for t in ts]
    for c in range(2):
        _emulated_all_to_all([s[c] for s in sends], [t.send_sizes for t in ts], [t.cols[c] for t in ts], [t.recv_sizes for t in ts])
    for t in ts:
        t.z_pass(0, 0)
    _emulated_all_to_all([t.dist_cols for t in ts], [t.recv_sizes for t in ts], [t.back for t in ts], [t.send_sizes for t in ts])
    one_sided = torch.cat([t.result() for t in ts]).cpu().numpy()
    assert np.array_equal(one_sided, edt.distance(dv, sampling).cpu().numpy())


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 4), (9, 17, 33), (12, 40, 70), (20, 33, 130), (5, 6, 1100), (40, 8, 9)])
@pytest.mark.parametrize("sampling", [(1.0, 1.0, 1.0), (0.09375, 0.31, 0.28)])
@pytest.mark.parametrize("kind", ["blobs", "noise", "sparse"])
def test_one_sweep_sdf_matches_scipy_and_the_two_transforms(eng, oracle, shape, sampling, kind):
    """t3d_sdf (both polarities in one sweep per axis, run-end zero-cost sites, float32-estimated take-over positions with an
    exact float64 check) against scipy's distance_transform_edt and against the two one-sided transforms."""
    from tomography_3d_reconstructor_b200 import edt
    rng = np.random.default_rng(sum(shape) + len(kind))
    if kind == "blobs":
        vol = random_blobs(rng, shape, 0.45, 1.5) if min(shape) > 1 else rng.random(shape) < 0.5
    elif kind == "noise":
        vol = rng.random(shape) < 0.5
    else:
        vol = rng.random(shape) < 0.02
    dv = dev_volume(eng, vol)
    got = edt.signed_distance(dv, sampling).cpu().numpy()
    two = edt.signed_distance_two_transforms(dv, sampling).cpu().numpy()
    ref = oracle.signed_distance(vol, sampling)
    if sampling == (1.0, 1.0, 1.0):
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))       # squared distances are exact integers
        assert np.array_equal(got.view(np.uint32), two.view(np.uint32))
    else:
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)
        assert np.allclose(got[fin], ref[fin], rtol=1e-6, atol=0) and np.allclose(got[fin], two[fin], rtol=1e-6, atol=0)
        assert np.array_equal(got[~fin], ref[~fin])

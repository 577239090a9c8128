"""Parity at BASELINE sizes (GPU box only).

C0 (104 x 512 x 512) runs through the oracle whole and is compared bit for bit through the host-facing entry point.
C1 (512 x 1024 x 1024) is too large for the CPU oracle to finish in seconds (its float64 field alone is 4.3 GB), so the
full-size GPU result is checked (a) against the oracle on z-slabs sampled from the top, the middle and the bottom of the
stack -- every operator of the path is local in z (gap fill 1, opening/closing 4, Gaussian 2, cube 1 planes), so the oracle
run on a slab extended by 10 slices reproduces the interior of the slab exactly -- and (b) through size-independent
properties: closed oriented 2-manifold, Euler characteristic 2, np.unique order, volumes against the analytic ellipsoid."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PHYS = (6.0, 143.1, 95.03)


def _rows_sorted(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort(tuple(a[:, k] for k in range(a.shape[1] - 1, -1, -1)))]


def _face_rows(v, f):
    """Each face as 9 floats (its three vertices' coordinates), rotated so the smallest vertex comes first."""
    tri = v[f].reshape(len(f), 3, 3).astype(np.float32).view(np.uint32).astype(np.int64)      # exact bit patterns (coordinates >= 0)
    n = len(f)
    fid = np.repeat(np.arange(n), 3)
    corner = np.tile(np.arange(3), n)
    flat = tri.reshape(3 * n, 3)
    order = np.lexsort((flat[:, 2], flat[:, 1], flat[:, 0], fid))     # per face: its lexicographically smallest vertex first
    k = corner[order[::3]]
    idx = (k[:, None] + np.arange(3)[None, :]) % 3
    rot = np.take_along_axis(tri, idx[:, :, None], axis=1).reshape(len(f), 9)
    return _rows_sorted(rot)


def test_config0_whole_stack_through_reconstruct_host(eng, oracle):
    from tomography_3d_reconstructor_b200 import pipeline
    Z, H, W = 104, 512, 512
    sides = (20, 64, 20)
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, 250:262, 240:270] = 0                      # a hole in the first slice: binary_fill_holes has work to do
    ref = oracle.reference_pipeline(u8, 200, sides, *PHYS)
    masks = [u8[z] for z in range(Z)]
    for _ in range(3):                               # staged (learns the sizes), fused, fused + graph replay
        out = pipeline.reconstruct_host(masks, 200, sides, *PHYS)
        assert np.array_equal(out["vertices"].view(np.uint32), ref["vertices"].view(np.uint32))
        assert np.array_equal(out["faces"], ref["faces"])
        assert out["voxel_volume_mm3"] == ref["voxel_volume"] and out["processed_voxel_volume_mm3"] == ref["processed_volume"]
        assert abs(out["mesh_volume_mm3"] - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]
        assert abs(out["surface_area_mm2"] - ref["surface_area"]) <= 1e-6 * ref["surface_area"]
        bb = out["bbox_index"]
        assert bb is not None


def test_config1_full_size_against_oracle_slabs_and_invariants(eng, oracle):
    import bench
    from tomography_3d_reconstructor_b200 import engine, pipeline
    Z, H, W = 512, 1024, 1024
    dev = torch.device("cuda", 0)
    sides = bench.side_counts(Z)
    masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
    pipeline._plans.clear(); pipeline._hints.clear()
    for _ in range(2):
        out = pipeline.reconstruct_fused(masks, 200, sides, *PHYS)
    v_d, f_d = out["mesh"].verts, out["mesh"].faces
    V, F = int(v_d.shape[0]), int(f_d.shape[0])
    # ---- (b) invariants of the whole mesh
    closed, euler = bench.mesh_topology(v_d, f_d)
    assert closed and euler == 2
    d = v_d[1:].double() - v_d[:-1].double()
    assert bool(((d[:, 0] > 0) | ((d[:, 0] == 0) & ((d[:, 1] > 0) | ((d[:, 1] == 0) & (d[:, 2] > 0))))).all())
    analytic_vox = 4 / 3 * np.pi * (0.42 * Z) * (0.33 * H) * (0.45 * W)
    assert abs(out["active_voxels"] / analytic_vox - 1) < 2e-3
    assert abs(out["mesh_volume_mm3"] / out["processed_voxel_volume_mm3"] - 1) < 5e-3
    assert out["mesh"].n_ambiguous == 0
    v = v_d.cpu().numpy()
    f = f_d.cpu().numpy()
    # ---- (a) oracle on sampled slabs
    depths = oracle.calculate_slice_depths(PHYS[0], *sides)
    mm_x, mm_y = PHYS[1] / W, PHYS[2] / H
    halo, n = 10, 6
    # z map of integer un-padded plane index k (what a vertex lying on plane k gets), see surface_extractor.py:82-113
    def zmap(k):
        t = np.array([[float(k), 0.0, 0.0]], dtype=np.float32)
        oracle.apply_variable_slice_depths(t, depths, True)
        return t[0, 0]
    for za in (44, Z // 2 - 3, Z - 50):              # the lower polar cap (the object spans planes 41..470), the equator, the upper cap
        zb = za + n
        a, b = za - halo, zb + halo
        u8 = oracle.ellipsoid_phantom_u8(Z, H, W, a, b)
        sub = oracle.smooth_voxel_data(oracle.create_voxel_data([u8[z] >= 200 for z in range(b - a)], True), 3, True)
        vol32 = oracle.scalar_field(sub, True, True)
        rv, rf, namb = oracle.marching_cubes(vol32, 0.5, z_base=a)     # vertex z from the global padded plane index
        assert namb == 0
        zi = rv[:, 0] - 1                             # un-padded plane coordinate of each vertex
        rv -= 1
        oracle.apply_variable_slice_depths(rv, depths, True)
        rv[:, 1] *= mm_y
        rv[:, 2] *= mm_x
        lo, hi = zmap(za), zmap(zb)
        keep_v = (rv[:, 0] >= lo) & (rv[:, 0] < hi)       # the same criterion on both sides: the mapped z value
        assert ((zi[keep_v] > za - 1) & (zi[keep_v] < zb + 1)).all()
        got_keep = (v[:, 0] >= lo) & (v[:, 0] < hi)
        ref_rows = _rows_sorted(np.unique(rv[keep_v], axis=0).view(np.uint32))
        got_rows = v[got_keep].view(np.uint32)          # the GPU list is already in np.unique order
        assert ref_rows.shape == got_rows.shape and np.array_equal(ref_rows, _rows_sorted(got_rows)), za
        assert np.array_equal(got_rows, _rows_sorted(got_rows))
        # faces with all three vertices inside the slab: the same triangles (same winding)
        keep_f = keep_v[rf].all(axis=1)
        ref_faces = _face_rows(rv, rf[keep_f])
        gf = f[got_keep[f].all(axis=1)]
        got_faces = _face_rows(v, gf)
        assert ref_faces.shape == got_faces.shape and np.array_equal(ref_faces, got_faces), za
        assert len(got_faces) > 1000
    # the smoothed voxel count of the whole stack against the oracle's on the middle slab, through the staged classes
    del masks
    pipeline._plans.clear()
    torch.cuda.empty_cache()

"""Pin the oracle against the UNMODIFIED reference modules, wherever those run without scikit-image.
Skipped when /root/reference is absent (the GPU box); the same checks are frozen in tests/golden/."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    import importlib.util
    import scipy.ndimage

    def load(name):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    vp, se, vc = load("voxel_processor"), load("surface_extractor"), load("volume_calculator")
    vp.SCIPY_AVAILABLE = True   # skimage absent => the import fallback cleared it although scipy is installed
    vp.ndimage = scipy.ndimage
    return vp, se, vc


def test_degraded_reference_behaviour_is_not_what_we_reproduce(ref):
    """without skimage the reference smooths nothing and extracts nothing."""
    vp, se, _ = ref
    x = np.zeros((4, 8, 8), bool)
    x[1:3, 2:6, 2:6] = True
    assert not vp.SKIMAGE_AVAILABLE and not se.SKIMAGE_AVAILABLE
    assert vp.VoxelProcessor().smooth_voxel_data(x) is x
    assert se.SurfaceExtractor().extract_manifold_surface(x, np.ones(4), 1.0, 1.0) is None


def test_close_volume_ends_and_volumes(ref, oracle):
    vp, _, vc = ref
    rng = np.random.default_rng(0)
    V = vc.VolumeCalculator()
    for _ in range(60):
        Z, H, W = rng.integers(1, 9), rng.integers(2, 12), rng.integers(2, 14)
        vol = rng.random((Z, H, W)) < rng.uniform(0.2, 0.8)
        masks = [vol[z] for z in range(Z)]
        with contextlib.redirect_stdout(io.StringIO()):
            P = vp.VoxelProcessor()
            got = P.create_voxel_data(masks, True, 1, max(Z - 2, 0), 1 if Z > 1 else 0)
            d = P.calculate_slice_depths(6.0)
        assert np.array_equal(oracle.create_voxel_data(masks, True), got)
        assert np.array_equal(oracle.close_volume_ends_stencil(vol), got)
        assert np.array_equal(oracle.calculate_slice_depths(6.0, P.side_0_count, P.side_1_count, P.side_2_count), d)
        assert oracle.calculate_voxel_volume_variable_depth(got, 0.3, 0.2, d) == V.calculate_voxel_volume_variable_depth(got, 0.3, 0.2, d)
        assert oracle.calculate_bounding_box_variable_depth(got, 0.3, 0.2, d) == V.calculate_bounding_box_variable_depth(got, 0.3, 0.2, d)
        if got.any():
            assert oracle.calculate_bounding_box(got, 0.3, 0.2, 0.1) == V.calculate_bounding_box(got, 0.3, 0.2, 0.1)
        for sub in (1, 2, 5):
            assert np.array_equal(oracle.generate_point_cloud(got, 0.3, 0.2, d, sub), P.generate_point_cloud(got, 0.3, 0.2, d, sub))


def test_surface_postprocessing(ref, oracle):
    _, se, _ = ref
    S = se.SurfaceExtractor()
    rng = np.random.default_rng(1)
    depths = oracle.calculate_slice_depths(6.0, 3, 10, 3)
    for pad in (True, False):
        z = (rng.random(5000) * 22 - 3).astype(np.float32)
        z[:50] = np.arange(50, dtype=np.float32) - 10
        a = np.zeros((len(z), 3), np.float32)
        a[:, 0] = z
        b = a.copy()
        S._apply_variable_slice_depths(a, depths, pad)
        oracle.apply_variable_slice_depths(b, depths, pad)
        assert np.array_equal(a, b)
    verts = (rng.integers(0, 30, (3000, 3)) * np.float32(0.37)).astype(np.float32)
    faces = rng.integers(0, 3000, (5000, 3)).astype(np.int32)
    rv, rf = S._ensure_manifold_mesh(verts, faces)
    ov, of = oracle.ensure_manifold_mesh(verts, faces)
    assert np.array_equal(rv, ov) and np.array_equal(rf, of) and of.dtype == rf.dtype
    small_f = of[:300]
    assert oracle.calculate_mesh_volume_literal(ov, small_f) == S.calculate_mesh_volume(ov, small_f)
    assert oracle.calculate_surface_area(ov, of) == S.calculate_surface_area(ov, of)
    assert np.array_equal(S._add_volume_padding(np.ones((2, 3, 4), bool)), np.pad(np.ones((2, 3, 4), bool), 1))

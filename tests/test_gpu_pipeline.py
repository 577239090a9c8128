"""GPU parity of the whole-job pipeline (what bench.py times) against the oracle's reference_pipeline."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(24, 64, 96), (40, 100, 130), (32, 128, 128)])
def test_reconstruct_vs_oracle(eng, oracle, shape):
    from tomography_3d_reconstructor_b200 import pipeline
    Z, H, W = shape
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, H // 2 - 2:H // 2 + 2, W // 2 - 3:W // 2 + 3] = 0
    sides = (Z // 8, Z - 2 * (Z // 8), Z // 8)
    res = pipeline.reconstruct(torch.from_numpy(u8).cuda(), 200, sides, 6.0, 143.1, 95.03)
    ref = oracle.reference_pipeline(u8, 200, sides, 6.0, 143.1, 95.03)
    mesh = res["mesh"]
    v, f = mesh.verts.cpu().numpy(), mesh.faces.cpu().numpy()
    assert v.shape == ref["vertices"].shape and f.shape == ref["faces"].shape
    assert np.allclose(v, ref["vertices"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(f, ref["faces"])
    assert res["voxel_volume_mm3"] == ref["voxel_volume"]                 # bit-exact (integer counts, same f64 order)
    assert res["processed_voxel_volume_mm3"] == ref["processed_volume"]
    assert abs(res["mesh_volume_mm3"] - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]
    assert abs(res["surface_area_mm2"] - ref["surface_area"]) <= 1e-6 * ref["surface_area"]
    z, y, x = np.where(ref["voxel_data"])
    assert res["bbox_index"] == (z.min(), z.max(), y.min(), y.max(), x.min(), x.max())
    assert res["active_voxels"] == int(ref["voxel_data"].sum())
    assert mesh.n_ambiguous == ref["n_ambiguous"] == 0


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("shape", [(40, 64, 96), (53, 70, 130)])
def test_sharded_slabs_stitch_to_the_single_gpu_mesh(eng, oracle, shape, world):
    """All ranks of a z-slab sharded step emulated on ONE GPU (the halo exchange is a device copy here): the
    concatenation of the per-rank canonical slabs must be the single-GPU mesh bit for bit."""
    from tomography_3d_reconstructor_b200 import pipeline, sharded
    Z, H, W = shape
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, H // 2 - 2:H // 2 + 2, W // 2 - 3:W // 2 + 3] = 0
    sides = (Z // 8, Z - 2 * (Z // 8), Z // 8)
    full = torch.from_numpy(u8).cuda()
    ref = pipeline.reconstruct(full, 200, sides, 6.0, 143.1, 95.03)
    ranges = [sharded.slab_range(Z, r, world) for r in range(world)]
    slabs = [sharded.slab_pack(full[a:b].contiguous(), Z, a, 200, world) for a, b in ranges]
    for r in range(world):                      # what exchange_halos does over NCCL
        s = slabs[r]
        if s.hl:
            lo = slabs[r - 1]
            s.ext[:s.hl] = lo.ext[lo.hl + lo.n - s.hl:lo.hl + lo.n]
        if s.hh:
            hi = slabs[r + 1]
            s.ext[s.hl + s.n:] = hi.ext[hi.hl:hi.hl + s.hh]
    local = [sharded.slab_local(s, sides, 6.0, 143.1, 95.03) for s in slabs]
    host = torch.stack(local).cpu()
    raw_counts = np.concatenate([s.cnt_raw.cpu().numpy() for s in slabs]).astype(np.int64)
    sm_counts = np.concatenate([s.cnt_sm.cpu().numpy() for s in slabs]).astype(np.int64)
    outs = [sharded.finalize(s, r, host, raw_counts, sm_counts, [a for a, _ in ranges], sides, 6.0, 143.1, 95.03)
            for r, s in enumerate(slabs)]
    assert all(o["stitch_consistent"] for o in outs)
    v = torch.cat([o["verts"] for o in outs]).cpu().numpy()
    f = torch.cat([o["faces"] for o in outs]).cpu().numpy()
    rv, rf = ref["mesh"].verts.cpu().numpy(), ref["mesh"].faces.cpu().numpy()
    assert np.array_equal(v.view(np.uint32), rv.view(np.uint32))
    assert np.array_equal(f, rf)
    o = outs[0]
    assert o["voxel_volume_mm3"] == ref["voxel_volume_mm3"]
    assert o["processed_voxel_volume_mm3"] == ref["processed_voxel_volume_mm3"]
    assert o["bbox_index"] == ref["bbox_index"] and o["active_voxels"] == ref["active_voxels"]
    assert abs(o["mesh_volume_mm3"] - ref["mesh_volume_mm3"]) <= 1e-9 * ref["mesh_volume_mm3"]
    assert abs(o["surface_area_mm2"] - ref["surface_area_mm2"]) <= 1e-9 * ref["surface_area_mm2"]
    assert o["total_vertices"] == len(rv) and o["total_faces"] == len(rf)


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("shape", [(40, 64, 96), (53, 70, 130), (72, 40, 128), (50, 37, 256)])
def test_fused_slab_path_stitches_to_the_single_gpu_mesh(eng, oracle, shape, world):
    """t3d_reconstruct_slab + t3d_slab_stitch_faces (the sharded step without host synchronisation), all ranks emulated on
    ONE GPU: device copies stand in for the NCCL halo exchange and the all-gather of the result blocks."""
    from tomography_3d_reconstructor_b200 import pipeline, sharded
    Z, H, W = shape
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, H // 2 - 2:H // 2 + 2, W // 2 - 3:W // 2 + 3] = 0
    sides = (Z // 8, Z - 2 * (Z // 8), Z // 8)
    full = torch.from_numpy(u8).cuda()
    ref = pipeline.reconstruct(full, 200, sides, 6.0, 143.1, 95.03)
    rm = ref["mesh"]
    caps = pipeline._caps_from(rm.n_active, *rm.n_raw, rm.n_z, rm.n_raw[0])   # generous: the whole mesh per rank
    ranges = [sharded.slab_range(Z, r, world) for r in range(world)]
    plans = [sharded.FusedSlabPlan(b - a, H, W, Z, a, 200, sides, 6.0, 143.1, 95.03, 3, True, caps, full.device, r, world)
             for r, (a, b) in enumerate(ranges)]
    for pl, (a, b) in zip(plans, ranges):
        pl.pack(full[a:b].contiguous())
    # slabs the one-pass pack kernel takes (W % 128 == 0, >= 16 slices) run in the pre-filled mode, the others in the plain one
    assert [pl.pre_active for pl in plans] == [W % 128 == 0 and (b - a) >= 16 for a, b in ranges]
    for r, s in enumerate(plans):
        if s.hl:
            lo = plans[r - 1]
            s.ext[:s.hl] = lo.ext[lo.hl + lo.n - s.hl:lo.hl + lo.n]
        if s.hh:
            hi = plans[r + 1]
            s.ext[s.hl + s.n:] = hi.ext[hi.hl:hi.hl + s.hh]
    for pl in plans:
        pl.compute()
    gathered = torch.stack([pl.res for pl in plans])
    for pl in plans:
        pl.gathered.copy_(gathered)
        pl.stitch()
    h = gathered.cpu().numpy()
    assert not h[:, pipeline.R_OVERFLOW].any() and not h[:, pipeline.R_UNVERIFIED].any()
    outs = [sharded.assemble(pl, h) for pl in plans]
    assert all(o["stitch_consistent"] for o in outs)
    v = torch.cat([o["verts"] for o in outs]).cpu().numpy()
    f = torch.cat([o["faces"] for o in outs]).cpu().numpy()
    rv, rf = rm.verts.cpu().numpy(), rm.faces.cpu().numpy()
    assert np.array_equal(v.view(np.uint32), rv.view(np.uint32))
    assert np.array_equal(f, rf)
    o = outs[-1]
    assert o["voxel_volume_mm3"] == ref["voxel_volume_mm3"]
    assert o["processed_voxel_volume_mm3"] == ref["processed_voxel_volume_mm3"]
    assert o["bbox_index"] == ref["bbox_index"] and o["active_voxels"] == ref["active_voxels"]
    assert abs(o["mesh_volume_mm3"] - ref["mesh_volume_mm3"]) <= 1e-9 * ref["mesh_volume_mm3"]
    assert abs(o["surface_area_mm2"] - ref["surface_area_mm2"]) <= 1e-9 * ref["surface_area_mm2"]
    assert o["total_vertices"] == len(rv) and o["total_faces"] == len(rf)


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_single_enqueue_path_matches_staged_path(eng, oracle, use_graph):
    """t3d_reconstruct (one enqueue, device-resident sizes, optionally replayed from a CUDA graph) against the staged
    path and the oracle, including capacity overflow -> fallback."""
    from tomography_3d_reconstructor_b200 import pipeline
    Z, H, W = 40, 96, 128
    sides = (5, 30, 5)
    args = (200, sides, 6.0, 143.1, 95.03)
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, 46:50, 60:68] = 0
    masks = torch.from_numpy(u8).cuda()
    pipeline._plans.clear(); pipeline._hints.clear()
    first = pipeline.reconstruct_fused(masks, *args, use_graph=use_graph)      # staged: learns the sizes
    ref = oracle.reference_pipeline(u8, 200, sides, 6.0, 143.1, 95.03)
    for rep in range(3):                                                      # fused (graph replayed on reps 2, 3)
        out = pipeline.reconstruct_fused(masks, *args, use_graph=use_graph)
        v, f = out["mesh"].verts.cpu().numpy(), out["mesh"].faces.cpu().numpy()
        assert np.array_equal(v, ref["vertices"]) and np.array_equal(f, ref["faces"])
        assert out["voxel_volume_mm3"] == ref["voxel_volume"] == first["voxel_volume_mm3"]
        assert out["processed_voxel_volume_mm3"] == ref["processed_volume"]
        assert abs(out["mesh_volume_mm3"] - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]
        assert abs(out["surface_area_mm2"] - ref["surface_area"]) <= 1e-6 * ref["surface_area"]
        assert out["bbox_index"] == first["bbox_index"] and out["active_voxels"] == first["active_voxels"]
    assert len(pipeline._plans) == 1
    # a much larger object in the same buffer overflows the capacities: detected on the device, staged fallback
    big = oracle.ellipsoid_phantom_u8(Z, H, W).copy()
    big[:, 4:-4, 4:-4] = 255
    big[::2, ::3, ::5] = 0
    masks.copy_(torch.from_numpy(big))
    out = pipeline.reconstruct_fused(masks, *args, use_graph=use_graph)
    ref2 = oracle.reference_pipeline(big, 200, sides, 6.0, 143.1, 95.03)
    assert np.array_equal(out["mesh"].verts.cpu().numpy(), ref2["vertices"])
    assert np.array_equal(out["mesh"].faces.cpu().numpy(), ref2["faces"])
    assert out["voxel_volume_mm3"] == ref2["voxel_volume"]


@pytest.mark.parametrize("case", ["blobs_touching_everything", "box_on_slice0", "noise", "plate", "no_depths_like"])
def test_fused_structured_vertex_order_on_awkward_volumes(eng, oracle, case):
    """The structured canonical ordering (t3d_mesh_canonicalize_structured_dev: y-edge ranking, one z-edge sort, clamp-group
    sort) must give exactly the mesh of the staged path on volumes that touch slice 0 / the image border, flat faces
    and noise; the capacities tune themselves (clamp group provisioned on demand)."""
    from conftest import random_blobs
    from tomography_3d_reconstructor_b200 import pipeline
    rng = np.random.default_rng(7)
    Z, H, W = 24, 70, 100
    if case == "blobs_touching_everything":
        occ = random_blobs(rng, (Z, H, W), 0.5, 2.0)
    elif case == "box_on_slice0":
        occ = np.zeros((Z, H, W), bool)
        occ[0:14, 10:50, 0:80] = True            # touches slice 0 and the x = 0 border: flat cap under slice 0
    elif case == "noise":
        occ = rng.random((Z, H, W)) < 0.55
    elif case == "plate":
        occ = np.zeros((Z, H, W), bool)
        occ[8:12, :, :] = True                   # spans the whole image: long y-edge / z-edge runs with equal keys
    else:
        occ = random_blobs(rng, (Z, H, W), 0.3, 3.0)
        occ[:2] = False
    u8 = (occ * 255).astype(np.uint8)
    masks = torch.from_numpy(u8).cuda()
    sides = (4, 16, 4)
    args = (200, sides, 6.0, 143.1, 95.03)
    pipeline._plans.clear(); pipeline._hints.clear(); pipeline._g0_caps.clear(); pipeline._generic_sort.clear()
    ref = pipeline.reconstruct(masks, *args)
    rv, rf = ref["mesh"].verts.cpu().numpy(), ref["mesh"].faces.cpu().numpy()
    for rep in range(4):          # 1: staged (learns), 2: fused, maybe re-tuned, 3-4: steady state
        out = pipeline.reconstruct_fused(masks, *args, use_graph=False)
        v, f = out["mesh"].verts.cpu().numpy(), out["mesh"].faces.cpu().numpy()
        assert np.array_equal(v.view(np.uint32), rv.view(np.uint32)) and np.array_equal(f, rf)
        assert out["voxel_volume_mm3"] == ref["voxel_volume_mm3"]
    assert len(pipeline._plans) == 1, "steady state must be the fused path"
    plan = next(iter(pipeline._plans.values()))
    if case in ("blobs_touching_everything", "box_on_slice0"):
        assert plan.caps[3] > 0 and plan.caps[4] > 0, "clamp group must be handled by the structured path, not a fallback"
    if case in ("plate", "no_depths_like"):
        assert plan.caps[3] > 0 and plan.caps[4] == 0


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sdf_pipeline_stitches_to_the_single_gpu_mesh(eng, oracle, world):
    """BASELINE configs 3/4 in miniature: smoothing on z-slabs, sharded exact signed distance (all-to-all transpose),
    marching cubes on the distance field with a one-plane float halo, ghost-plane stitching -- all ranks emulated on ONE
    GPU; the stitched mesh and the distance field must equal pipeline.reconstruct_sdf on the whole stack bit for bit."""
    from tomography_3d_reconstructor_b200 import edt, pipeline, sharded
    from test_gpu_edt import _emulated_all_to_all
    Z, H, W = 44, 60, 90
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, H // 2 - 2:H // 2 + 2, W // 2 - 3:W // 2 + 3] = 0
    sides = (5, 34, 5)
    phys = (6.0, 143.1, 95.03)
    full = torch.from_numpy(u8).cuda()
    ref = pipeline.reconstruct_sdf(full, 200, sides, *phys)
    rv, rf = ref["mesh"].verts.cpu().numpy(), ref["mesh"].faces.cpu().numpy()
    assert len(rf) > 1000
    depths = pipeline.slice_depths(phys[0], *sides)
    samp = pipeline.sdf_sampling(depths, phys[2] / H, phys[1] / W)
    ranges = [sharded.slab_range(Z, r, world) for r in range(world)]
    slabs = [sharded.slab_pack(full[a:b].contiguous(), Z, a, 200, world) for a, b in ranges]
    for r, s in enumerate(slabs):
        if s.hl:
            lo = slabs[r - 1]
            s.ext[:s.hl] = lo.ext[lo.hl + lo.n - s.hl:lo.hl + lo.n]
        if s.hh:
            hi = slabs[r + 1]
            s.ext[s.hl + s.n:] = hi.ext[hi.hl:hi.hl + s.hh]
    sms = [sharded.sdf_slab_smooth(s) for s in slabs]
    ts = [edt.SlabTransform(sm.bits, Z, a, H, W, samp, r, world) for r, (sm, (a, b)) in enumerate(zip(sms, ranges))]
    sends = [t.sdf_xy_pass() for t in ts]
    for c in range(2):
        _emulated_all_to_all([x[c] for x in sends], [t.send_sizes for t in ts], [t.cols[c] for t in ts], [t.recv_sizes for t in ts])
    for t in ts:
        t.sdf_z_pass()
    _emulated_all_to_all([t.dist_cols for t in ts], [t.recv_sizes for t in ts], [t.back for t in ts], [t.send_sizes for t in ts])
    sdfs = [t.result() for t in ts]
    assert torch.equal(torch.cat(sdfs), ref["sdf"])
    local = [sharded.sdf_slab_surface(s, sdfs[r], sdfs[r + 1][0] if r + 1 < world else None, sides, *phys)
             for r, s in enumerate(slabs)]
    host = torch.stack(local).cpu()
    raw_counts = np.concatenate([s.cnt_raw.cpu().numpy() for s in slabs]).astype(np.int64)
    sm_counts = np.concatenate([s.cnt_sm.cpu().numpy() for s in slabs]).astype(np.int64)
    outs = [sharded.finalize(s, r, host, raw_counts, sm_counts, [a for a, _ in ranges], sides, *phys) for r, s in enumerate(slabs)]
    assert all(o["stitch_consistent"] for o in outs)
    v = torch.cat([o["verts"] for o in outs]).cpu().numpy()
    f = torch.cat([o["faces"] for o in outs]).cpu().numpy()
    assert np.array_equal(v.view(np.uint32), rv.view(np.uint32)) and np.array_equal(f, rf)
    o = outs[0]
    assert o["voxel_volume_mm3"] == ref["voxel_volume_mm3"] and o["processed_voxel_volume_mm3"] == ref["processed_voxel_volume_mm3"]
    assert abs(o["mesh_volume_mm3"] - ref["mesh_volume_mm3"]) <= 1e-9 * ref["mesh_volume_mm3"]
    # the distance-field surface encloses about the same volume as the voxel count (sanity of the whole SDF path)
    assert abs(ref["mesh_volume_mm3"] - ref["processed_voxel_volume_mm3"]) < 0.05 * ref["processed_voxel_volume_mm3"]


def test_batch_of_independent_phantoms_matches_the_oracle(eng, oracle):
    """BASELINE configs[2] in miniature: independent random ellipsoid phantoms through batch.reconstruct_batch (one shared
    plan + CUDA graph, items copied into its input buffer) against the oracle, item by item; two 'ranks' cover the batch."""
    from tomography_3d_reconstructor_b200 import batch, pipeline
    n, count = 40, 5
    radii, centres = batch.phantom_params(count, n)
    dev = torch.device("cuda", 0)
    stacks = [batch.phantom_u8(n, radii[i], centres[i], dev) for i in range(count)]
    sides = (5, 30, 5)
    phys = (6.0, 143.1, 95.03)
    pipeline._plans.clear(); pipeline._hints.clear()
    got = {}
    for rank in range(2):
        got.update(batch.reconstruct_batch(stacks, 200, sides, *phys, rank=rank, world=2))
    assert sorted(got) == list(range(count))
    for i in range(count):
        ref = oracle.reference_pipeline(stacks[i].cpu().numpy(), 200, sides, *phys)
        assert np.array_equal(got[i]["vertices"], ref["vertices"]) and np.array_equal(got[i]["faces"], ref["faces"])
        assert got[i]["voxel_volume_mm3"] == ref["voxel_volume"] and got[i]["processed_voxel_volume_mm3"] == ref["processed_volume"]
        assert abs(got[i]["mesh_volume_mm3"] - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]


def test_many_volumes_in_flight_match_the_one_at_a_time_batch(eng, oracle):
    """batch.reconstruct_volumes (several plans / streams / captured graphs in flight) against batch.reconstruct_batch and
    the oracle; the second call is the steady state, a volume larger than any seen before falls back and is re-learned."""
    from tomography_3d_reconstructor_b200 import batch, pipeline
    n, count = 40, 7
    radii, centres = batch.phantom_params(count, n, seed=99)
    dev = torch.device("cuda", 0)
    stacks = {i: batch.phantom_u8(n, radii[i], centres[i], dev) for i in range(count)}
    sides = (5, 30, 5)
    phys = (6.0, 143.1, 95.03)
    batch._batch_caps.clear(); batch._batch_slots.clear()
    small = {i: stacks[i] for i in range(count - 1)}
    for rep in range(3):          # 1: learns the capacities (staged), 2-3: in flight
        got = batch.reconstruct_volumes(small, 200, sides, *phys, keep_mesh=True, slots=3)
        assert sorted(got) == sorted(small)
        for i in small:
            ref = oracle.reference_pipeline(stacks[i].cpu().numpy(), 200, sides, *phys)
            assert np.array_equal(got[i]["vertices"], ref["vertices"]) and np.array_equal(got[i]["faces"], ref["faces"]), (rep, i)
            assert got[i]["voxel_volume_mm3"] == ref["voxel_volume"] and got[i]["processed_voxel_volume_mm3"] == ref["processed_volume"]
            assert abs(got[i]["mesh_volume_mm3"] - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]
    # a much larger object than the learned capacities allow: staged fallback for that item, capacities grow
    big = torch.zeros((n, n, n), dtype=torch.uint8, device=dev)
    big[1:-1, 1:-1, 1:-1] = 255
    big[::2, ::2, 5:35:2] = 0
    mixed = dict(small)
    mixed[count] = big
    got = batch.reconstruct_volumes(mixed, 200, sides, *phys, keep_mesh=True, slots=3)
    ref = oracle.reference_pipeline(big.cpu().numpy(), 200, sides, *phys)
    assert np.array_equal(got[count]["vertices"], ref["vertices"]) and np.array_equal(got[count]["faces"], ref["faces"])


@pytest.mark.parametrize("shape", [(1, 9, 33), (2, 5, 7), (3, 64, 31), (4, 3, 257), (5, 33, 130), (9, 40, 1)])
def test_fused_path_on_degenerate_shapes(eng, oracle, shape):
    """Tiny / ragged stacks (one or two slices, rows narrower or slightly wider than a machine word, W = 1) through the
    single-enqueue path against the oracle, bit for bit; also when nothing is left to mesh."""
    from conftest import random_blobs
    from tomography_3d_reconstructor_b200 import pipeline
    rng = np.random.default_rng(sum(shape))
    Z, H, W = shape
    occ = random_blobs(rng, shape, 0.6, 1.0) if min(shape) > 1 else rng.random(shape) < 0.7
    u8 = (occ * 255).astype(np.uint8)
    sides = (0, Z, 0)
    args = (200, sides, 6.0, 143.1, 95.03)
    masks = torch.from_numpy(u8).cuda()
    pipeline._plans.clear(); pipeline._hints.clear(); pipeline._g0_caps.clear(); pipeline._generic_sort.clear()
    ref = oracle.reference_pipeline(u8, 200, sides, 6.0, 143.1, 95.03)
    for rep in range(3):
        if ref.get("vertices") is None:      # the oracle found nothing to mesh (reference: extract -> None)
            with pytest.raises((RuntimeError, ValueError)):
                pipeline.reconstruct_fused(masks, *args, use_graph=False)
            continue
        out = pipeline.reconstruct_fused(masks, *args, use_graph=(rep == 2))
        v, f = out["mesh"].verts.cpu().numpy(), out["mesh"].faces.cpu().numpy()
        assert np.array_equal(v.view(np.uint32), ref["vertices"].view(np.uint32)) and np.array_equal(f, ref["faces"]), rep
        assert out["voxel_volume_mm3"] == ref["voxel_volume"] and out["processed_voxel_volume_mm3"] == ref["processed_volume"]


def test_second_device_in_one_process(eng, oracle):
    """State that lives per device (constant-memory tables, function attributes, side streams) is initialised per device: the
    whole path on cuda:1 after cuda:0 in ONE process (needs two GPUs; skipped otherwise)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tomography_3d_reconstructor_b200 import pipeline
    Z, H, W = 24, 64, 96
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, 30:34, 40:44] = 0
    sides = (3, 18, 3)
    ref = oracle.reference_pipeline(u8, 200, sides, 6.0, 143.1, 95.03)
    try:
        for d in (0, 1, 0):
            torch.cuda.set_device(d)
            masks = torch.from_numpy(u8).to("cuda:%d" % d)
            for _ in range(3):      # staged, fused, fused + graph
                out = pipeline.reconstruct_fused(masks, 200, sides, 6.0, 143.1, 95.03)
                assert out["mesh"].verts.device.index == d
                assert np.array_equal(out["mesh"].verts.cpu().numpy(), ref["vertices"]) and np.array_equal(out["mesh"].faces.cpu().numpy(), ref["faces"])
                assert out["voxel_volume_mm3"] == ref["voxel_volume"]
    finally:
        torch.cuda.set_device(0)

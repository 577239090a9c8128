"""Host-side logic of the z-slab sharded path, on CPU: world_size-2 gloo runs of the halo exchange / gathers, and the
stitching rule (per-slab canonical lists concatenate to the global np.unique mesh) replayed with the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_slab_ranges_partition_the_stack():
    from tomography_3d_reconstructor_b200 import sharded
    for Z in (16, 17, 53, 512, 4096):
        for world in (1, 2, 3, 4, 8):
            r = [sharded.slab_range(Z, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == Z
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert sharded.owned_padded_planes(100, 0, 25) == (0, 26)
    assert sharded.owned_padded_planes(100, 25, 50) == (26, 51)
    assert sharded.owned_padded_planes(100, 75, 100) == (76, 102)
    assert sharded.stitch_offsets([(10, 3, 0), (9, 2, 3), (5, 0, 2)]) == ([0, 7, 14], True)
    assert sharded.stitch_offsets([(10, 3, 0), (9, 2, 4), (5, 0, 2)])[1] is False


def test_per_slab_canonical_meshes_concatenate(oracle):
    """The rule sharded.py relies on, checked with numpy only: np.unique order is z-major, so slab r's sorted vertices
    end with slab r+1's leading (first-plane) vertices and global ids are base[r] + local id."""
    from tomography_3d_reconstructor_b200 import sharded, engine
    Z, H, W = 40, 48, 64
    vol = oracle.smooth_voxel_data(oracle.create_voxel_data([m for m in (oracle.ellipsoid_phantom_u8(Z, H, W) >= 200)]), 3, True)
    depths = oracle.calculate_slice_depths(6.0, 5, 30, 5)
    mm_y, mm_x = 95.03 / H, 143.1 / W
    gv, gf = oracle.extract_manifold_surface(vol, depths, mm_y, mm_x)
    field = oracle.scalar_field(vol, True, True)
    rv, rf, _ = oracle.marching_cubes(field, 0.5)             # raw global mesh, padded index coordinates
    layer = np.floor(rv[rf, 0].min(axis=1)).astype(int)       # cube layer of every face
    tv = rv.copy()
    tv -= 1
    oracle.apply_variable_slice_depths(tv, depths, True)
    tv[:, 1] *= mm_y
    tv[:, 2] *= mm_x
    cum, adj = engine.z_map_arrays(depths, True)
    for world in (2, 3):
        per_rank, meshes = [], []
        for r in range(world):
            z0, z1 = sharded.slab_range(Z, r, world)
            a, b = sharded.owned_padded_planes(Z, z0, z1)
            f = rf[(layer >= a) & (layer < b)]                 # the faces rank r emits (its cube layers) ...
            used, inv = np.unique(f, return_inverse=True)      # ... and the vertices they use (own + ghost plane b)
            uv, uf = oracle.ensure_manifold_mesh(tv[used], inv.reshape(-1, 3))
            ghost = int((uv[:, 0] == sharded.z_map_value(b - 1, cum, adj)).sum()) if z1 < Z else 0
            lead = int((uv[:, 0] == sharded.z_map_value(a - 1, cum, adj)).sum()) if z0 > 0 else 0
            per_rank.append((len(uv), ghost, lead))
            meshes.append((uv, uf))
        bases, ok = sharded.stitch_offsets(per_rank)
        assert ok
        sv = np.concatenate([uv[:n - g] for (uv, _), (n, g, _) in zip(meshes, per_rank)])
        sf = np.concatenate([uf + bases[r] for r, (_, uf) in enumerate(meshes)])
        assert np.array_equal(sv, gv) and np.array_equal(sf, gf)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tomography_3d_reconstructor_b200 import sharded
        Z, H, wpr = 40, 6, 4
        glob = torch.arange(Z * H * wpr, dtype=torch.int32).reshape(Z, H, wpr)     # "packed planes" with unique values
        z0, z1 = sharded.slab_range(Z, rank, world)
        n = z1 - z0
        hl, hh = (sharded.HALO if z0 > 0 else 0), (sharded.HALO if z1 < Z else 0)
        ext = torch.full((hl + n + hh, H, wpr), -1, dtype=torch.int32)
        ext[hl:hl + n] = glob[z0:z1]
        sharded.exchange_halos(ext, hl, n, hh, rank, world)
        ok_halo = bool(torch.equal(ext, glob[z0 - hl:z1 + hh]))
        # the small all-gather + stitch arithmetic
        local = torch.tensor([100 + 10 * rank, 7, 0, 3 if rank + 1 < world else 0, 3 if rank else 0], dtype=torch.int64)
        got = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(got, local)
        per_rank = [(int(g[0]), int(g[3]), int(g[4])) for g in got]
        bases, ok = sharded.stitch_offsets(per_rank)
        # gather of per-rank mesh slabs
        res = {"verts": torch.full((2 + rank, 3), float(rank)), "faces": torch.full((3 + rank, 3), rank, dtype=torch.int64)}
        m = sharded.gather_mesh(res, 0)
        ok_gather = True
        if rank == 0:
            v, f = m
            ok_gather = v.shape[0] == sum(2 + r for r in range(world)) and f.shape[0] == sum(3 + r for r in range(world)) \
                and float(v[-1, 0]) == world - 1 and int(f[-1, 0]) == world - 1
        # the EDT's all-to-all transpose z-slabs -> y-slabs and back (edt.py), on CPU tensors
        from tomography_3d_reconstructor_b200 import edt
        Zt, Ht, Wt = 11, 7, 5
        g16 = (torch.arange(Zt * Ht * Wt, dtype=torch.int32) % 30000).to(torch.int16).reshape(Zt, Ht, Wt)
        zr, yr, send_sizes, recv_sizes = edt.transpose_plan(Zt, Ht, Wt, rank, world)
        a, b = zr[rank]
        ya, yb = yr[rank]
        cols = torch.empty(Zt * (yb - ya) * Wt, dtype=torch.int16)
        edt._exchange(edt.pack_rows(g16[a:b], yr), send_sizes, cols, recv_sizes, rank, world, None)
        ok_fwd = bool(torch.equal(cols.view(Zt, yb - ya, Wt), g16[:, ya:yb, :]))
        back = torch.empty((b - a) * Ht * Wt, dtype=torch.int16)
        edt._exchange(cols, recv_sizes, back, send_sizes, rank, world, None)
        ok_back = bool(torch.equal(edt.unpack_rows(back, b - a, Ht, Wt, yr), g16[a:b]))
        q.put((rank, ok_halo, bases, ok, ok_gather and ok_fwd and ok_back))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_halo_exchange_and_gathers_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, bases, ok, ok_gather in out:
        assert ok_halo and ok and ok_gather
        assert bases == [0, 97]


def test_balanced_ranges_tile_the_stack_with_equal_cost():
    import numpy as np
    from tomography_3d_reconstructor_b200 import sharded
    Z, world = 4096, 8
    z = np.arange(Z)
    verts = 4000.0 * np.sqrt(np.clip(1 - ((z - Z / 2) / (0.42 * Z)) ** 2, 0, None))     # an ellipsoid's vertices per slice
    cost = sharded.slice_cost(verts)
    ranges = sharded.balanced_ranges(cost, world, sharded.HALO)
    assert ranges[0][0] == 0 and ranges[-1][1] == Z and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    assert all(b - a >= sharded.HALO for a, b in ranges)
    sums = np.array([cost[a:b].sum() for a, b in ranges])
    assert sums.max() / sums.mean() < 1.01                       # equal slices would give 1.19 on this profile
    eq = np.array([cost[a:b].sum() for a, b in [sharded.slab_range(Z, r, world) for r in range(world)]])
    assert eq.max() / eq.mean() > 1.1
    # the end slabs (polar caps, little surface) get more slices than the equatorial ones
    assert ranges[0][1] - ranges[0][0] > ranges[world // 2][1] - ranges[world // 2][0]
    # registering the partition redirects slab_range for exactly this (Z, world)
    try:
        sharded.set_partition(Z, world, ranges)
        assert [sharded.slab_range(Z, r, world) for r in range(world)] == ranges
        assert sharded.slab_range(Z, 0, 4) == (0, Z // 4)
        with pytest.raises(ValueError):
            sharded.set_partition(Z, world, ranges[:-1])
    finally:
        sharded.set_partition(Z, world, None)
    assert sharded.slab_range(Z, 1, world) == (Z // 8, Z // 4)
    # degenerate: everything in one slice still leaves every rank its minimum
    spike = np.zeros(64); spike[10] = 1.0
    r2 = sharded.balanced_ranges(spike, 4, 8)
    assert all(b - a >= 8 for a, b in r2) and r2[-1][1] == 64
    with pytest.raises(ValueError):
        sharded.balanced_ranges(np.ones(20), 4, 8)

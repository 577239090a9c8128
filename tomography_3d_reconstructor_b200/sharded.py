"""z-slab sharded reconstruction: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

The reference is single-process; slices are independent images, so the stack shards naturally into
contiguous z-slabs.  Per step each rank

  1. packs its own slices (rank 0 / the last rank also fill the holes of global slice 0 / Z-1);
  2. exchanges HALO = 8 bit-planes with each z-neighbour in ONE grouped send/recv (1 MB per side at 1024x1024):
     1 plane for the z gap fill + 4 for the opening/closing + 3 for the Gaussian radius and the cube's upper corners.
     Gap fill and morphology are simply recomputed on the halo planes, so no second exchange is needed;
  3. marches the cube layers it owns.  A grid edge belongs to the rank that owns the plane of its lower corner; the
     x/y-edge vertices of the first plane of the NEXT rank are emitted as ghosts so that the top cube layer is closed;
  4. canonicalises its mesh locally.  The canonical order (np.unique: z, then y, then x) is z-major, so a rank's
     sorted list ends with exactly the vertices the next rank's list starts with (same bits, same order): global
     ids are `base[rank] + local id` with base = the cross-rank exclusive scan of the owned unique-vertex counts, the
     ghost tail is dropped, and the concatenation over ranks IS the single-GPU mesh, bit for bit;
  5. all-gathers the per-slice voxel counts (host float64 sum in the reference's order) and all-reduces bbox / mesh
     volume / area partials.

No data-path collective other than the neighbour halo exchange and KB-sized count gathers.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import engine, pipeline

HALO = 8          # bit-planes exchanged per side
SURF_HALO = 3     # smoothed planes the surface stage needs beyond the owned ones


_partitions: Dict = {}      # (Z, world) -> [(z0, z1)] set by set_partition(); every rank must set the same one


def slab_range(Z: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice range [z0, z1) of `rank`: the registered partition of this (Z, world), else equal slices."""
    part = _partitions.get((int(Z), int(world)))
    if part is not None:
        return part[rank]
    base, rem = divmod(Z, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def balanced_ranges(per_slice_cost, world: int, min_slices: int = 8) -> List[Tuple[int, int]]:
    """Contiguous z-slabs of (nearly) equal summed cost.  Equal slice counts give equal voxel work but not equal surface
    work: the slabs that hold the polar caps of an object carry less mesh than the ones through its equator, and the
    step ends with a collective, so every rank waits for the heaviest slab.  per_slice_cost[z] = estimated cost of slice z
    (e.g. slice_cost()).  Every slab gets at least `min_slices` slices (the halo width)."""
    c = np.asarray(per_slice_cost, dtype=np.float64)
    Z = len(c)
    if Z < world * min_slices:
        raise ValueError("stack too thin for %d slabs of at least %d slices" % (world, min_slices))
    cum = np.concatenate([[0.0], np.cumsum(c)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        z = int(np.searchsorted(cum, target))
        if z > 0 and abs(cum[z - 1] - target) <= abs(cum[min(z, Z)] - target):
            z -= 1
        z = max(z, cuts[-1] + min_slices)
        z = min(z, Z - (world - r) * min_slices)
        cuts.append(z)
    cuts.append(Z)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def set_partition(Z: int, world: int, ranges) -> None:
    """Register the z-slab partition of a (Z, world) job (None: back to equal slices).  Collective by convention: all ranks
    must register the same ranges before the next step; plans built for another partition are dropped."""
    key = (int(Z), int(world))
    if ranges is None:
        _partitions.pop(key, None)
    else:
        ranges = [(int(a), int(b)) for a, b in ranges]
        if len(ranges) != world or ranges[0][0] != 0 or ranges[-1][1] != Z or any(ranges[r][1] != ranges[r + 1][0] for r in range(world - 1)):
            raise ValueError("ranges must tile [0, Z) in rank order")
        _partitions[key] = ranges
    _slab_plans.clear()
    _slab_hints.clear()


def slice_cost(vertices_per_slice, slice_equivalent_vertices: float = 2900.0) -> np.ndarray:
    """Cost model of one slice for balanced_ranges(): the volume passes cost the same for every slice (about as much as
    ~2900 mesh vertices at 1024x1024: 0.35 ms / 512 slices against 0.5 ms / 2.06 M vertices), the surface passes scale
    with the vertices the slice contributes."""
    return slice_equivalent_vertices + np.asarray(vertices_per_slice, dtype=np.float64)


def vertices_per_slice(res: Dict, Zg: int, add_padding: bool = True) -> np.ndarray:
    """Vertices of this rank's slab of the stitched mesh (result dict of reconstruct*) per global slice index, as a length-Zg
    histogram: a vertex with z between the map values of planes k and k+1 belongs to slice k (learning step only)."""
    depths = res["slice_depths"]
    cum, adj = engine.z_map_arrays(depths, add_padding)
    knots = np.array([z_map_value(k, cum, adj) for k in range(Zg + 1)], dtype=np.float64)
    vz = res["verts"][:, 0].double()
    idx = torch.bucketize(vz, torch.from_numpy(knots).to(vz.device), right=True) - 1
    idx = idx.clamp_(0, Zg - 1)
    return torch.bincount(idx, minlength=Zg).cpu().numpy().astype(np.float64)


def owned_padded_planes(Zg: int, z0: int, z1: int, pad: int = 1) -> Tuple[int, int]:
    """Planes [a, b) of the padded grid (Zg + 2*pad planes) whose lower-corner edges / cube layers a slab owns."""
    a = 0 if z0 == 0 else z0 + pad
    b = Zg + 2 * pad if z1 == Zg else z1 + pad
    return a, b


z_map_value = engine.z_map_value


def exchange_halos(ext: torch.Tensor, hl: int, n: int, hh: int, rank: int, world: int, group=None) -> None:
    """Fill ext[:hl] from rank-1 and ext[hl+n:] from rank+1; send our first / last HALO planes the other way."""
    ops = []
    if hl:
        ops.append(dist.P2POp(dist.irecv, ext[:hl], rank - 1, group))
        ops.append(dist.P2POp(dist.isend, ext[hl:hl + hl].contiguous(), rank - 1, group))
    if hh:
        ops.append(dist.P2POp(dist.irecv, ext[hl + n:], rank + 1, group))
        ops.append(dist.P2POp(dist.isend, ext[hl + n - hh:hl + n].contiguous(), rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def stitch_offsets(per_rank: List[Tuple[int, int, int]]) -> Tuple[List[int], bool]:
    """per_rank[r] = (unique vertices incl. ghost tail, ghost tail length, leading vertices on the first owned plane).
    Returns (global id base per rank, consistent?).  Consistent = every ghost tail is exactly the next rank's lead."""
    bases, run, ok = [], 0, True
    for r, (v, g, lead) in enumerate(per_rank):
        bases.append(run)
        run += v - g
        if r + 1 < len(per_rank):
            ok = ok and (g == per_rank[r + 1][2])
        else:
            ok = ok and g == 0
    return bases, ok


class Slab:
    """Per-rank state between the phases of a sharded step."""
    __slots__ = ("ext", "hl", "n", "hh", "z0", "z1", "Zg", "H", "W", "dev", "mesh", "cnt_raw", "cnt_sm", "local")


def slab_pack(masks_u8: torch.Tensor, Zg: int, z0: int, threshold: int, world: int) -> Slab:
    """Phase 1: pack own slices into the middle of the halo-extended buffer, fill the holes of global end slices."""
    L = engine._L()
    p, st = engine._p, engine._stream
    s = Slab()
    s.n, s.H, s.W = (int(v) for v in masks_u8.shape)
    s.z0, s.z1, s.Zg, s.dev = z0, z0 + s.n, Zg, masks_u8.device
    if world > 1 and s.n < HALO:
        raise ValueError("z-slabs must be at least %d slices thick" % HALO)
    s.hl, s.hh = (HALO if z0 > 0 else 0), (HALO if s.z1 < Zg else 0)
    wpr = engine.words_per_row(s.W)
    s.ext = torch.empty((s.hl + s.n + s.hh, s.H, wpr), dtype=torch.int32, device=s.dev)
    engine.check(L.t3d_pack_masks(p(masks_u8), s.n, s.H, s.W, int(threshold), p(s.ext[s.hl]), st()), "t3d_pack_masks")
    ends = ([s.hl] if z0 == 0 else []) + ([s.hl + s.n - 1] if s.z1 == Zg and not (z0 == 0 and s.n == 1) else [])
    if ends:
        scratch = torch.empty(int(L.t3d_fill_holes_scratch_bytes(1, s.H, s.W)) // 4, dtype=torch.int32, device=s.dev)
        for e in ends:
            engine.check(L.t3d_fill_holes_2d(p(s.ext[e]), 1, 0, s.H, s.W, p(scratch), st()), "t3d_fill_holes_2d")
    return s


def slab_local(s: Slab, side_counts, total_depth_mm: float, x_length_mm: float, y_length_mm: float, iterations: int = 3,
               add_padding: bool = True, mark=None) -> torch.Tensor:
    """Phases 3-4 (after the halo exchange): gap fill + smoothing on the extended buffer, surface of the owned cube
    layers, local canonical mesh.  Returns the int64 vector this rank contributes to the small all-gather:
    [V' incl. ghosts, F', unverified-order flag, ghost tail, lead, signed volume bits, area bits, bbox(6, local z)]."""
    mark = mark or (lambda _n: None)
    L = engine._L()
    p, st = engine._p, engine._stream
    hl, n, hh, H, W, dev = s.hl, s.n, s.hh, s.H, s.W, s.dev
    mm_x, mm_y = x_length_mm / W, y_length_mm / H
    depths = pipeline.slice_depths(total_depth_mm, *side_counts)
    Zx = hl + n + hh
    gf = torch.empty_like(s.ext)
    cnt_raw = torch.empty(Zx, dtype=torch.int64, device=dev)
    engine.check(L.t3d_gap_fill(p(s.ext), p(gf), None, None, Zx, H, W, p(cnt_raw), st()), "t3d_gap_fill")
    raw = engine.DeviceVolume(gf, Zx, H, W, cnt_raw)
    smx = engine.smooth(raw, iterations, True)
    mark("smooth")
    s.cnt_raw, s.cnt_sm = cnt_raw[hl:hl + n], smx.counts_tensor()[hl:hl + n]
    sl, sh = min(SURF_HALO, hl), min(SURF_HALO, hh)
    loc = engine.DeviceVolume(smx.bits[hl - sl:hl + n + sh], sl + n + sh, H, W)
    pad = 1 if add_padding else 0
    z_offset = s.z0 - sl
    a, b = owned_padded_planes(s.Zg, s.z0, s.z1, pad)
    bbox_t = engine.DeviceVolume(gf[hl:hl + n], n, H, W, s.cnt_raw).bbox_tensor()
    try:
        mesh = engine.extract_surface(loc, depths, mm_y, mm_x, True, add_padding, canonical="async", mark=mark,
                                      z_begin=a - z_offset, z_end=b - z_offset, z_offset=z_offset)
        raw_verts, raw_faces = mesh.raw
        meas = engine.mesh_measure_async(raw_verts, raw_faces)
        counts_mesh = mesh.counts_dev
    except RuntimeError:       # this slab holds no surface
        mesh, meas = None, torch.zeros(2, dtype=torch.float64, device=dev)
        counts_mesh = torch.zeros(3, dtype=torch.int64, device=dev)
    mark("measure")
    s.mesh = mesh
    # ghost tail / lead: canonical vertices lying exactly on the next rank's first plane / on our first plane
    # (the vertex transform subtracts 1 from the padded plane index before the z map, surface_extractor.py:57-60)
    cum, adj = engine.z_map_arrays(depths, add_padding)
    zero = torch.zeros((), dtype=torch.int64, device=dev)
    n_ghost = n_lead = zero
    if mesh is not None:
        vz = mesh._verts[:, 0]
        valid = torch.arange(vz.shape[0], device=dev) < counts_mesh[0]
        if s.z1 < s.Zg:
            n_ghost = ((vz == float(z_map_value(b - pad, cum, adj))) & valid).sum()
        if s.z0 > 0:
            n_lead = ((vz == float(z_map_value(a - pad, cum, adj))) & valid).sum()
    s.local = torch.cat([counts_mesh, n_ghost.reshape(1), n_lead.reshape(1), meas.view(torch.int64), bbox_t.to(torch.int64)])
    return s.local


def finalize(s: Slab, rank: int, host, raw_counts: np.ndarray, sm_counts: np.ndarray, slab_starts: List[int],
             side_counts, total_depth_mm: float, x_length_mm: float, y_length_mm: float, stitched=None, depths=None,
             vol_weights=None) -> Dict:
    """Phase 5: host[r] = the vector of slab_local() of every rank (int64, torch or numpy); counts = global per-slice
    voxel counts.  stitched = (capacity-sized canonical verts, faces already carrying global ids) from the fused path.
    Plain numpy / Python on purpose: this runs on the host between two steps, with the GPU idle."""
    mm_x, mm_y = x_length_mm / s.W, y_length_mm / s.H
    if depths is None:
        depths = pipeline.slice_depths(total_depth_mm, *side_counts)
    h = host.numpy() if isinstance(host, torch.Tensor) else np.asarray(host)
    rows = h.tolist()
    per_rank = [(row[0], row[3], row[4]) for row in rows]
    bases, consistent = stitch_offsets(per_rank)
    consistent = consistent and not any(row[2] for row in rows)   # a rank whose fast ordering failed needs the general path
    meas_all = np.ascontiguousarray(h[:, 5:7]).view(np.float64)
    signed_volume, area = float(meas_all[:, 0].sum()), float(meas_all[:, 1].sum())
    bbox = None
    nonempty = [r for r, row in enumerate(rows) if row[8] >= 0]
    if nonempty:
        bbox = (min(rows[r][7] + slab_starts[r] for r in nonempty), max(rows[r][8] + slab_starts[r] for r in nonempty),
                min(rows[r][9] for r in nonempty), max(rows[r][10] for r in nonempty),
                min(rows[r][11] for r in nonempty), max(rows[r][12] for r in nonempty))
    v_own = per_rank[rank][0] - per_rank[rank][1]
    n_faces = rows[rank][1]
    if stitched is not None:
        verts_own, faces_global = stitched[0][:v_own], stitched[1][:n_faces]
    elif s.mesh is not None:
        s.mesh.set_sizes(per_rank[rank][0], n_faces, 0)
        verts_own = s.mesh._verts[:v_own]
        faces_global = s.mesh._faces + bases[rank]
    else:
        verts_own = torch.empty((0, 3), dtype=torch.float32, device=s.dev)
        faces_global = torch.empty((0, 3), dtype=torch.int64, device=s.dev)
    total_v, total_f = sum(v - g for v, g, _ in per_rank), sum(row[1] for row in rows)
    return {
        "verts": verts_own, "faces": faces_global, "vertex_base": bases[rank], "stitch_consistent": consistent,
        "total_vertices": total_v, "total_faces": total_f,
        "voxel_volume_mm3": pipeline.variable_depth_volume(raw_counts, mm_x, mm_y, depths, vol_weights),
        "processed_voxel_volume_mm3": pipeline.variable_depth_volume(sm_counts, mm_x, mm_y, depths, vol_weights),
        "mesh_volume_mm3": abs(signed_volume), "surface_area_mm2": area, "bbox_index": bbox,
        "active_voxels": int(raw_counts.sum()), "slice_depths": depths,
        "mesh": _MeshView(verts_own, faces_global, total_v, total_f), "local_mesh": s.mesh,
    }


def reconstruct(masks_u8: torch.Tensor, Zg: int, z0: int, threshold: int, side_counts, total_depth_mm: float,
                x_length_mm: float, y_length_mm: float, iterations: int = 3, add_padding: bool = True,
                mark: Optional[Callable[[str], None]] = None, group=None) -> Dict:
    """masks_u8: this rank's slices [z0, z0+n) of the global (Zg,H,W) uint8 stack, on its GPU."""
    mark = mark or (lambda _n: None)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s = slab_pack(masks_u8, Zg, z0, threshold, world)
    mark("pack_close")
    if world > 1:
        exchange_halos(s.ext, s.hl, s.n, s.hh, rank, world, group)
    mark("halo")
    local = slab_local(s, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations, add_padding, mark)
    sizes = [slab_range(Zg, r, world) for r in range(world)]
    if world > 1:
        # ONE small all-gather: [local result vector | raw per-slice counts | smoothed per-slice counts] (padded)
        nmax = max(e - b for b, e in sizes)
        nl = int(local.numel())
        msg = torch.zeros(nl + 2 * nmax, dtype=torch.int64, device=s.dev)
        msg[:nl] = local
        msg[nl:nl + s.n] = s.cnt_raw
        msg[nl + nmax:nl + nmax + s.n] = s.cnt_sm
        allmsg = torch.empty((world, nl + 2 * nmax), dtype=torch.int64, device=s.dev)
        dist.all_gather_into_tensor(allmsg, msg, group=group)
        hm = allmsg.cpu()
        host = hm[:, :nl].contiguous()
        hc = hm[:, nl:].numpy()
        raw_counts = np.concatenate([hc[r, :e - b] for r, (b, e) in enumerate(sizes)]).astype(np.int64)
        sm_counts = np.concatenate([hc[r, nmax:nmax + e - b] for r, (b, e) in enumerate(sizes)]).astype(np.int64)
    else:
        host = local.cpu().reshape(1, -1)
        raw_counts = s.cnt_raw.cpu().numpy().astype(np.int64)
        sm_counts = s.cnt_sm.cpu().numpy().astype(np.int64)
    mark("stats")
    return finalize(s, rank, host, raw_counts, sm_counts, [b for b, _ in sizes], side_counts, total_depth_mm,
                    x_length_mm, y_length_mm)


# ----------------------------------------------------------------------------------------------------------------
# fused sharded step: pack -> halo exchange -> ONE t3d_reconstruct_slab enqueue -> all-gather of the result blocks ->
# face stitching -> one device->host copy.  No host synchronisation before the end of the step.
# ----------------------------------------------------------------------------------------------------------------
class FusedSlabPlan:
    """Per-rank buffers for one (slab shape, parameters); capacities from a previous staged step of the same input.
    The returned mesh slab lives in the plan's buffers: valid until the next run()."""

    def __init__(self, n, H, W, Zg, z0, threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations,
                 add_padding, caps, device, rank, world, group=None):
        L = engine._L()
        self.n, self.H, self.W, self.Zg, self.z0, self.z1 = int(n), int(H), int(W), int(Zg), int(z0), int(z0) + int(n)
        self.rank, self.world, self.group, self.dev = rank, world, group, device
        if world > 1 and self.n < HALO:
            raise ValueError("z-slabs must be at least %d slices thick" % HALO)
        self.threshold, self.add_padding = int(threshold), bool(add_padding)
        self.side_counts, self.phys = tuple(side_counts), (float(total_depth_mm), float(x_length_mm), float(y_length_mm))
        self.mm_x, self.mm_y = x_length_mm / W, y_length_mm / H
        self.depths = pipeline.slice_depths(total_depth_mm, *side_counts)
        self.vol_weights = pipeline.volume_weights(self.mm_x, self.mm_y, self.depths)
        stages = engine.morph_stages(iterations, True)
        self.n_stages = len(stages)
        self.erode_mask = sum(1 << k for k, er in enumerate(stages) if er)
        self.caps = tuple(int(c) for c in caps)
        if len(self.caps) != 5:
            raise ValueError("caps = (active words, vertices, faces, z-edge vertices, clamp group)")
        self.hl, self.hh = (HALO if self.z0 > 0 else 0), (HALO if self.z1 < Zg else 0)
        pad = 1 if add_padding else 0
        sl = min(SURF_HALO, self.hl)
        self.z_offset = self.z0 - sl
        a, b = owned_padded_planes(Zg, self.z0, self.z1, pad)
        self.z_begin, self.z_end = a - self.z_offset, b - self.z_offset
        cum, adj = engine.z_map_arrays(self.depths, add_padding)
        self.n_cum = len(cum)
        self.cum_d = torch.from_numpy(cum).to(device) if self.n_cum else None
        self.adj_d = torch.from_numpy(adj).to(device) if self.n_cum else None
        # the vertex transform subtracts `pad` from the padded plane index before the z map (surface_extractor.py:57-60)
        self.want_ghost, self.want_lead = int(self.z1 < Zg), int(self.z0 > 0)
        self.z_ghost = float(z_map_value(b - pad, cum, adj)) if self.want_ghost else 0.0
        self.z_lead = float(z_map_value(a - pad, cum, adj)) if self.want_lead else 0.0
        n_surf = min(SURF_HALO, self.hl) + self.n + min(SURF_HALO, self.hh) + 2 * pad
        self.zkey_bits = engine.zkey_bits(self.depths, add_padding, n_surf, self.z_offset, 1)
        wpr = engine.words_per_row(W)
        Zx = self.hl + self.n + self.hh
        self.ext = torch.zeros((Zx, H, wpr), dtype=torch.int32, device=device)
        self.fill = torch.empty(int(L.t3d_fill_holes_scratch_bytes(2, H, W)) // 4 + 1, dtype=torch.int32, device=device)
        nbytes = int(L.t3d_reconstruct_slab_workspace_bytes(self.hl, self.n, self.hh, H, W, pad, self.n_stages, *self.caps))
        self.ws = torch.empty(nbytes // 8 + 1, dtype=torch.int64, device=device)
        self.verts = torch.empty((self.caps[1], 3), dtype=torch.float32, device=device)
        self.faces = torch.empty((self.caps[2], 3), dtype=torch.int64, device=device)
        self.sizes = [slab_range(Zg, r, world) for r in range(world)]
        nmax = max(e - b_ for b_, e in self.sizes)
        self.stride = pipeline.R_COUNTS + 2 * (nmax + 2 * HALO)
        self.res = torch.zeros(self.stride, dtype=torch.int64, device=device)
        self.gathered = torch.zeros((world, self.stride), dtype=torch.int64, device=device)
        self.host = torch.zeros((world, self.stride), dtype=torch.int64, pin_memory=True)
        self.host_np = self.host.numpy()
        self.graph, self.graph_ptr = None, None
        # where every rank's own per-slice counts sit in its result block (raw, then smoothed)
        self.raw_spans, self.sm_spans = [], []
        for b_, e in self.sizes:
            hl_r, hh_r = (HALO if b_ > 0 else 0), (HALO if e < Zg else 0)
            zx_r = hl_r + (e - b_) + hh_r
            c0 = pipeline.R_COUNTS
            self.raw_spans.append((c0 + hl_r, c0 + hl_r + (e - b_)))
            self.sm_spans.append((c0 + zx_r + hl_r, c0 + zx_r + hl_r + (e - b_)))
        self.slab_starts = [b_ for b_, _ in self.sizes]
        # positions of the global per-slice counts (row 0: raw, row 1: smoothed) in the flattened gathered blocks
        self.count_index = np.stack([
            np.concatenate([r * self.stride + np.arange(a_, b_) for r, (a_, b_) in enumerate(spans)])
            for spans in (self.raw_spans, self.sm_spans)])
        self.last_view, self.last_view_key = None, None
        # pre-filled mode (see pack()): interior planes of the gap-filled grid written by the pack kernel itself
        self.grid, self.pre, self.pre_active = None, None, False
        # compute() joins the library's side stream when pack() started the hole filling of a global end slice there
        self.join_fill = self.z0 == 0 or self.z1 == self.Zg

    def pack(self, masks_u8: torch.Tensor) -> None:
        """Own slices -> planes [hl, hl+n) of the extended buffer; the holes of the global end slices are filled on the
        library's side stream (joined by compute()).  Slabs the one-pass kernel takes (t3d_slab_pack_gap_ok) go through the
        pre-filled mode: the interior own planes are thresholded, gap-filled and counted in ONE pass over the masks into
        self.grid, only the 8 planes at either end are packed raw for the halo exchange (include/t3d.h, t3d_slab_pack)."""
        p = engine._p
        L = engine._L()
        self.pre_active = bool(L.t3d_slab_pack_gap_ok(p(masks_u8), self.n, self.H, self.W, self.threshold))
        if self.pre_active and self.grid is None:
            Zx = self.hl + self.n + self.hh
            self.grid = torch.empty((Zx, self.H, engine.words_per_row(self.W)), dtype=torch.int32, device=self.dev)
            self.pre = torch.empty(Zx + 3, dtype=torch.int64, device=self.dev)
        engine.check(L.t3d_slab_pack(p(masks_u8), self.n, self.H, self.W, self.threshold, self.hl, self.hh,
                                     int(self.z0 == 0), int(self.z1 == self.Zg), p(self.ext), p(self.fill),
                                     p(self.grid) if self.pre_active else None, p(self.pre) if self.pre_active else None,
                                     engine._stream()), "t3d_slab_pack")

    def compute(self) -> None:
        """Everything between the halo exchange and the result gather: one t3d_reconstruct_slab enqueue."""
        p = engine._p
        engine.check(engine._L().t3d_reconstruct_slab(
            p(self.ext), self.hl, self.n, self.hh, self.H, self.W, self.n_stages, self.erode_mask, 1 if self.add_padding else 0,
            self.z_begin, self.z_end, self.z_offset, self.want_ghost, self.z_ghost, self.want_lead, self.z_lead,
            int(self.join_fill), engine._W3_C,
            p(self.cum_d), p(self.adj_d), self.n_cum, float(self.mm_y), float(self.mm_x), 0, self.caps[0], self.caps[1],
            self.caps[2], self.caps[3], self.caps[4], self.zkey_bits, p(self.verts), p(self.faces), p(self.res), p(self.ws),
            p(self.grid) if self.pre_active else None, p(self.pre) if self.pre_active else None, engine._stream()), "t3d_reconstruct_slab")

    def stitch(self) -> None:
        """Local face ids -> ids in the stitched mesh, from the gathered result blocks (device side)."""
        engine.check(engine._L().t3d_slab_stitch_faces(engine._p(self.faces), self.caps[2], engine._p(self.gathered), self.stride,
                                                      self.rank, engine._stream()), "t3d_slab_stitch_faces")

    def enqueue(self, masks_u8: torch.Tensor) -> None:
        self.pack(masks_u8)
        if self.world > 1:
            exchange_halos(self.ext, self.hl, self.n, self.hh, self.rank, self.world, self.group)
        self.compute()
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.res, group=self.group)
            self.stitch()
        else:
            self.gathered[0].copy_(self.res)

    def capture(self, masks_u8: torch.Tensor) -> None:
        """Record the whole step (NCCL halo exchange and result all-gather included) for this input buffer into a CUDA
        graph; every rank must capture, and replay, in the same step."""
        self.enqueue(masks_u8)
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=engine.capture_stream()):
            self.enqueue(masks_u8)
            self.host.copy_(self.gathered, non_blocking=True)     # the read-back of the result blocks is a node of the graph
        self.graph, self.graph_ptr = g, masks_u8.data_ptr()

    def run(self, masks_u8: torch.Tensor, use_graph: bool = False) -> np.ndarray:
        if use_graph and self.graph is not None and self.graph_ptr == masks_u8.data_ptr():
            self.graph.replay()
        else:
            self.enqueue(masks_u8)
            self.host.copy_(self.gathered, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.host_np


_slab_plans: Dict = {}
_slab_hints: Dict = {}
_slab_retuned: Dict = {}


def reconstruct_fused(masks_u8: torch.Tensor, Zg: int, z0: int, threshold: int, side_counts, total_depth_mm: float,
                      x_length_mm: float, y_length_mm: float, iterations: int = 3, add_padding: bool = True, group=None,
                      use_graph: bool = False, collective_capture: bool = False) -> Dict:
    """Same contract and results as reconstruct(); the step is enqueued without any host synchronisation (use_graph:
    replayed from a CUDA graph that includes the NCCL operations; all ranks must pass the same value).

    A captured graph is bound to the input buffer it was captured on, and (re)capturing runs the step's collectives, so
    with use_graph EVERY rank has to (re)capture in the same call.  That holds when all ranks pass persistent buffers (what
    reconstruct_host and bench.py do) or all pass fresh ones.  A caller that cannot promise it sets collective_capture=True:
    the ranks then agree on "somebody needs a capture" with one small all-reduce per call (a host synchronisation, ~50 us:
    not the default).

    The first call for a given (slab, parameters) runs the staged path to learn the mesh sizes; if any rank reports a
    capacity overflow, an unverifiable fast ordering or an empty slab, every rank re-runs the staged path."""
    R = pipeline
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n, H, W = (int(v) for v in masks_u8.shape)
    key = (n, H, W, int(Zg), int(z0), int(threshold), tuple(side_counts), float(total_depth_mm), float(x_length_mm),
           float(y_length_mm), int(iterations), bool(add_padding), masks_u8.device.index, world)

    def staged():
        out = reconstruct(masks_u8, Zg, z0, threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations,
                          add_padding, group=group)
        m = out.get("local_mesh")
        _slab_hints[key] = pipeline._caps_from(m.n_active, *m.n_raw, m.n_z) if m is not None else (4096, 4096, 4096, 4096, 0)
        return out

    if key not in _slab_hints:
        return staged()
    caps = pipeline._tuned_caps(("slab",) + key, _slab_hints[key])
    plan = _slab_plans.get(key)
    if plan is None:      # plans are (re)built by all ranks in the same call: building one runs the step's collectives once more
        plan = FusedSlabPlan(n, H, W, Zg, z0, threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations,
                             add_padding, caps, masks_u8.device, rank, world, group)
        _slab_plans[key] = plan
    if use_graph:
        need = plan.graph_ptr != masks_u8.data_ptr()
        if collective_capture and world > 1:
            flag = torch.tensor([1 if need else 0], dtype=torch.int32, device=masks_u8.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            need = bool(int(flag.item()))
        if need:
            plan.capture(masks_u8)
    h = plan.run(masks_u8, use_graph)
    me = h[rank]
    bad = h[:, R.R_UNVERIFIED] != 0
    if bad.any() and not (h[:, R.R_OVERFLOW] != 0).any() and not _slab_retuned.get(key, 0) >= 2:
        # every rank takes this branch together (the decision only uses gathered data); a rank whose own ordering was
        # not verified provisions its clamp-group sort or switches to the generic 64-bit sort
        _slab_retuned[key] = _slab_retuned.get(key, 0) + 1
        if bad[rank] and plan.caps[3]:
            pipeline._retune(("slab",) + key, int(me[R.R_NG0]), plan.caps[4])
        _slab_plans.pop(key, None)
        return reconstruct_fused(masks_u8, Zg, z0, threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations,
                                 add_padding, group, use_graph, collective_capture)
    if (h[:, R.R_OVERFLOW] != 0).any() or bad.any() or (h[:, R.R_NT] == 0).any():
        _slab_plans.pop(key, None)
        _slab_hints.pop(key, None)
        return staged()
    _slab_hints[key] = tuple(max(a, b) for a, b in zip(_slab_hints[key], pipeline._caps_from(
        int(me[R.R_NACTIVE]), int(me[R.R_VRAW]), int(me[R.R_NT]), int(me[R.R_NZ]), int(me[R.R_NG0]))))
    return assemble(plan, h)


R_NGHOST, R_NLEAD = 18, 19

_device_inputs: Dict = {}


def reconstruct_host(masks_host: np.ndarray, Zg: int, z0: int, threshold: int, side_counts, total_depth_mm: float,
                     x_length_mm: float, y_length_mm: float, iterations: int = 3, add_padding: bool = True, group=None,
                     use_graph: bool = True) -> Dict:
    """The sharded step on HOST data: this rank's slices [z0, z0+n) as a numpy uint8 (or bool: use threshold=1) array
    (n,H,W), ideally in pinned memory -> upload -> reconstruct_fused -> this rank's slab of the stitched mesh as numpy
    arrays in out["vertices_host"], out["faces_host"] (global vertex ids; concatenating the ranks' arrays in rank order
    gives the single-GPU mesh)."""
    a = np.ascontiguousarray(masks_host)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    if a.dtype != np.uint8 or a.ndim != 3:
        raise ValueError("masks_host must be a (n,H,W) uint8 or bool array")
    dev = engine._require_cuda()
    key = (a.shape, dev.index if dev.index is not None else torch.cuda.current_device())
    buf = _device_inputs.get(key)
    if buf is None:                                  # persistent input buffer: keeps the captured graph valid
        buf = _device_inputs[key] = torch.empty(a.shape, dtype=torch.uint8, device=dev)
    buf.copy_(torch.from_numpy(a), non_blocking=True)
    out = reconstruct_fused(buf, Zg, z0, threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations,
                            add_padding, group, use_graph)
    v, f = out["verts"].contiguous(), out["faces"].contiguous()
    hv = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
    hf = torch.empty(f.shape, dtype=f.dtype, pin_memory=True)
    hv.copy_(v, non_blocking=True)
    hf.copy_(f, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    out["vertices_host"], out["faces_host"] = hv.numpy(), hf.numpy()
    return out


# ----------------------------------------------------------------------------------------------------------------
# bit-packed host input (additive API): 1 bit per voxel over PCIe instead of 1 byte
# ----------------------------------------------------------------------------------------------------------------
def pack_bits_host(mask_images) -> np.ndarray:
    """(Z,H,W) bool / 0-1 masks (array or list) -> the library's packed layout on the host: uint32 (Z,H,words_per_row(W)), bit i of
    word w = voxel x = 32*w + i, zero beyond W.  For callers whose masks already live packed (or that can pack while decoding)."""
    a = engine._as_stack(mask_images)
    Z, H, W = a.shape
    wpr = engine.words_per_row(W)
    out = np.zeros((Z, H, wpr * 4), dtype=np.uint8)
    out[:, :, :(W + 7) // 8] = np.packbits(a != 0, axis=-1, bitorder="little")
    return out.view("<u4")


_bits_plans: Dict = {}


def reconstruct_host_bits(bits_host: np.ndarray, W: int, side_counts, total_depth_mm: float, x_length_mm: float, y_length_mm: float,
                          iterations: int = 3, add_padding: bool = True, use_graph: bool = True) -> Dict:
    """pipeline.reconstruct_host for masks that are already bit-packed on the host (pack_bits_host layout; ideally pinned):
    the upload is W/8 bytes per row instead of W -- the host->device copy is what bounds reconstruct_host (537 MB of the
    14 ms at 512 x 1024 x 1024).  Same results (out["vertices"], out["faces"], volumes)."""
    b = np.ascontiguousarray(bits_host)
    if b.dtype.itemsize != 4 or b.ndim != 3 or b.shape[2] != engine.words_per_row(W):
        raise ValueError("bits_host must be (Z, H, words_per_row(W)) 32-bit words")
    Z, H = int(b.shape[0]), int(b.shape[1])
    dev = engine._require_cuda()
    L = engine._L()
    key = (Z, H, int(W), tuple(side_counts), float(total_depth_mm), float(x_length_mm), float(y_length_mm), int(iterations),
           bool(add_padding), torch.cuda.current_device())
    src = torch.from_numpy(b.view(np.int32))
    st = _bits_plans.get(key)
    if st is None:
        # learning step: unpack on the device and run the staged path once for the mesh sizes
        bits_d = src.to(dev)
        u8 = torch.empty((Z, H, W), dtype=torch.uint8, device=dev)
        engine.check(L.t3d_unpack_bits(engine._p(bits_d), Z, H, W, engine._p(u8), engine._stream()), "t3d_unpack_bits")
        out = pipeline.reconstruct(u8, 1, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations, True, add_padding)
        m = out["mesh"]
        caps = pipeline._caps_from(m.n_active, *m.n_raw, m.n_z)
        plan = FusedSlabPlan(Z, H, W, Z, 0, 1, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations, add_padding,
                             pipeline._tuned_caps(("bits",) + key, caps), dev, 0, 1)
        plan.join_fill = False          # the end slices are hole-filled on the main stream here (no t3d_slab_pack)
        st = _bits_plans[key] = {"plan": plan, "graph": None}
        out["vertices"], out["faces"] = engine.download(m.verts), engine.download(m.faces)
        return out
    plan = st["plan"]

    def enqueue():
        p = engine._p
        if Z >= 1:
            ends = [0] if Z == 1 else [0, Z - 1]
            for e in ends:
                engine.check(L.t3d_fill_holes_2d(p(plan.ext[e]), 1, 0, H, W, p(plan.fill), engine._stream()), "t3d_fill_holes_2d")
        plan.compute()
        plan.gathered[0].copy_(plan.res)
        plan.host.copy_(plan.gathered, non_blocking=True)

    plan.ext.copy_(src, non_blocking=True)
    if use_graph:
        if st["graph"] is None:
            enqueue()
            torch.cuda.current_stream().synchronize()
            plan.ext.copy_(src, non_blocking=True)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=engine.capture_stream()):
                enqueue()
            st["graph"] = g
        st["graph"].replay()
    else:
        enqueue()
    torch.cuda.current_stream().synchronize()
    h = plan.host_np
    R = pipeline
    if h[0, R.R_OVERFLOW] or h[0, R.R_UNVERIFIED] or h[0, R.R_NT] == 0:
        _bits_plans.pop(key, None)          # sizes changed beyond the margin (or an unverifiable ordering): learn again
        return reconstruct_host_bits(bits_host, W, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations, add_padding,
                                     use_graph)
    out = assemble(plan, h)
    v, f = out["verts"].contiguous(), out["faces"].contiguous()
    hv = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
    hf = torch.empty(f.shape, dtype=f.dtype, pin_memory=True)
    hv.copy_(v, non_blocking=True)
    hf.copy_(f, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    out["vertices"], out["faces"] = hv.numpy(), hf.numpy()
    return out


def assemble(plan: FusedSlabPlan, h: np.ndarray) -> Dict:
    """Result dict of reconstruct() from the gathered result blocks h (world x stride int64, host).  Same arithmetic as
    finalize(), written for speed (a few numpy calls, the rest plain Python): it runs between two steps, GPU idle."""
    R = pipeline
    rank = plan.rank
    rows = h[:, :R.R_COUNTS].tolist()
    per_rank = [(row[R.R_VCANON], row[R_NGHOST], row[R_NLEAD]) for row in rows]
    bases, consistent = stitch_offsets(per_rank)
    consistent = consistent and not any(row[R.R_UNVERIFIED] for row in rows)
    signed_volume, area = np.ascontiguousarray(h[:, R.R_VOLUME:R.R_VOLUME + 2]).view(np.float64).sum(axis=0).tolist()
    bbs = np.ascontiguousarray(h[:, R.R_BBOX:R.R_BBOX + 3]).view(np.int32).tolist()
    nonempty = [r for r, bb in enumerate(bbs) if bb[1] >= 0]
    bbox = None
    if nonempty:
        st = plan.slab_starts
        bbox = (min(bbs[r][0] + st[r] for r in nonempty), max(bbs[r][1] + st[r] for r in nonempty),
                min(bbs[r][2] for r in nonempty), max(bbs[r][3] for r in nonempty),
                min(bbs[r][4] for r in nonempty), max(bbs[r][5] for r in nonempty))
    # global per-slice counts, raw and smoothed, in one gather: (2, Zg)
    counts2 = h.ravel()[plan.count_index]
    n = min(counts2.shape[1], len(plan.depths))
    vols = np.cumsum(counts2[:, :n].astype(np.float64) * plan.vol_weights[:n], axis=1)[:, -1].tolist() if n else [0.0, 0.0]
    v_own, n_faces = per_rank[rank][0] - per_rank[rank][1], rows[rank][R.R_FCANON]
    if plan.last_view is None or plan.last_view_key != (v_own, n_faces):
        plan.last_view = (plan.verts[:v_own], plan.faces[:n_faces])      # views of the plan's buffers: valid until the next run
        plan.last_view_key = (v_own, n_faces)
    verts_own, faces_global = plan.last_view
    total_v, total_f = sum(v - g for v, g, _ in per_rank), sum(row[R.R_FCANON] for row in rows)
    mesh = _MeshView(verts_own, faces_global, total_v, total_f)
    mesh.n_ambiguous = sum(row[R.R_NAMBIGUOUS] for row in rows)
    return {
        "verts": verts_own, "faces": faces_global, "vertex_base": bases[rank], "stitch_consistent": consistent,
        "total_vertices": total_v, "total_faces": total_f,
        "voxel_volume_mm3": vols[0], "processed_voxel_volume_mm3": vols[1],
        "mesh_volume_mm3": abs(signed_volume), "surface_area_mm2": area, "bbox_index": bbox,
        "active_voxels": int(counts2[0].sum()), "slice_depths": plan.depths,
        "mesh": mesh, "local_mesh": None, "n_ambiguous": mesh.n_ambiguous,
    }


# ----------------------------------------------------------------------------------------------------------------
# SDF variant (BASELINE configs 3/4): smoothing on z-slabs as above, exact signed distance through the all-to-all
# transpose of edt.py, marching cubes on the distance field with ONE float32 halo plane from the next rank, the same
# ghost-plane stitching.  Written as phases so that tests can run every rank of a job in one process.
# ----------------------------------------------------------------------------------------------------------------
def sdf_slab_smooth(s: Slab, iterations: int = 3):
    """After slab_pack + the halo exchange: gap fill and smoothing on the extended buffer.  Returns the own smoothed
    planes as a DeviceVolume and sets the per-slice counts / bbox tensor on the slab."""
    L = engine._L()
    p, st = engine._p, engine._stream
    hl, n, hh, H, W = s.hl, s.n, s.hh, s.H, s.W
    Zx = hl + n + hh
    gf = torch.empty_like(s.ext)
    cnt_raw = torch.empty(Zx, dtype=torch.int64, device=s.dev)
    engine.check(L.t3d_gap_fill(p(s.ext), p(gf), None, None, Zx, H, W, p(cnt_raw), st()), "t3d_gap_fill")
    smx = engine.smooth(engine.DeviceVolume(gf, Zx, H, W, cnt_raw), iterations, True)
    s.cnt_raw, s.cnt_sm = cnt_raw[hl:hl + n], smx.counts_tensor()[hl:hl + n]
    s.local = engine.DeviceVolume(gf[hl:hl + n], n, H, W, s.cnt_raw).bbox_tensor()     # kept for sdf_slab_surface
    return engine.DeviceVolume(smx.bits[hl:hl + n].contiguous(), n, H, W)


def exchange_sdf_plane(sdf_own: torch.Tensor, rank: int, world: int, group=None) -> Optional[torch.Tensor]:
    """First own plane -> previous rank; returns the next rank's first plane (None on the last rank)."""
    nxt = torch.empty_like(sdf_own[0]) if rank + 1 < world else None
    ops = []
    if nxt is not None:
        ops.append(dist.P2POp(dist.irecv, nxt, rank + 1, group))
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, sdf_own[0].contiguous(), rank - 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return nxt


def sdf_slab_surface(s: Slab, sdf_own: torch.Tensor, sdf_next: Optional[torch.Tensor], side_counts, total_depth_mm: float,
                     x_length_mm: float, y_length_mm: float, level: float = 0.0) -> torch.Tensor:
    """Marching cubes on the own planes of the distance field (+ the next rank's first plane as ghost plane), local
    canonical mesh.  Returns the vector of slab_local()."""
    n, H, W, dev = s.n, s.H, s.W, s.dev
    mm_x, mm_y = x_length_mm / W, y_length_mm / H
    depths = pipeline.slice_depths(total_depth_mm, *side_counts)
    field = torch.cat([sdf_own, sdf_next[None]]) if sdf_next is not None else sdf_own
    bbox_t = s.local
    try:
        mesh = engine.extract_surface(None, depths, mm_y, mm_x, False, False, canonical="async", z_begin=0,
                                      z_end=n if sdf_next is not None else -1, z_offset=s.z0, field=field.contiguous(), level=level)
        meas = engine.mesh_measure_async(*mesh.raw)
        counts_mesh = mesh.counts_dev
    except RuntimeError:       # this slab holds no surface
        mesh, meas = None, torch.zeros(2, dtype=torch.float64, device=dev)
        counts_mesh = torch.zeros(3, dtype=torch.int64, device=dev)
    s.mesh = mesh
    cum, adj = engine.z_map_arrays(depths, False)
    zero = torch.zeros((), dtype=torch.int64, device=dev)
    n_ghost = n_lead = zero
    if mesh is not None:
        vz = mesh._verts[:, 0]
        valid = torch.arange(vz.shape[0], device=dev) < counts_mesh[0]
        if s.z1 < s.Zg:
            n_ghost = ((vz == float(z_map_value(s.z1, cum, adj))) & valid).sum()
        if s.z0 > 0:
            n_lead = ((vz == float(z_map_value(s.z0, cum, adj))) & valid).sum()
    s.local = torch.cat([counts_mesh, n_ghost.reshape(1), n_lead.reshape(1), meas.view(torch.int64), bbox_t.to(torch.int64)])
    return s.local


def reconstruct_sdf(masks_u8: torch.Tensor, Zg: int, z0: int, threshold: int, side_counts, total_depth_mm: float,
                    x_length_mm: float, y_length_mm: float, iterations: int = 3, level: float = 0.0, sampling=None,
                    group=None) -> Dict:
    """z-slab sharded pipeline.reconstruct_sdf: every rank passes its slices [z0, z0+n) of the global stack and gets its
    slab of the stitched mesh (+ "sdf": its slices of the distance field)."""
    from . import edt
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s = slab_pack(masks_u8, Zg, z0, threshold, world)
    if world > 1:
        exchange_halos(s.ext, s.hl, s.n, s.hh, rank, world, group)
    sm = sdf_slab_smooth(s, iterations)
    depths = pipeline.slice_depths(total_depth_mm, *side_counts)
    samp = sampling or pipeline.sdf_sampling(depths, y_length_mm / s.H, x_length_mm / s.W)
    sdf = edt.signed_distance_sharded(sm.bits, Zg, z0, s.H, s.W, samp, group) if world > 1 else edt.signed_distance(sm, samp)
    nxt = exchange_sdf_plane(sdf, rank, world, group) if world > 1 else None
    local = sdf_slab_surface(s, sdf, nxt, side_counts, total_depth_mm, x_length_mm, y_length_mm, level)
    sizes = [slab_range(Zg, r, world) for r in range(world)]
    nmax = max(e - b for b, e in sizes)
    nl = int(local.numel())
    msg = torch.zeros(nl + 2 * nmax, dtype=torch.int64, device=s.dev)
    msg[:nl] = local
    msg[nl:nl + s.n] = s.cnt_raw
    msg[nl + nmax:nl + nmax + s.n] = s.cnt_sm
    if world > 1:
        allmsg = torch.empty((world, nl + 2 * nmax), dtype=torch.int64, device=s.dev)
        dist.all_gather_into_tensor(allmsg, msg, group=group)
    else:
        allmsg = msg[None]
    hm = allmsg.cpu()
    host = hm[:, :nl].contiguous()
    hc = hm[:, nl:].numpy()
    raw_counts = np.concatenate([hc[r, :e - b] for r, (b, e) in enumerate(sizes)]).astype(np.int64)
    sm_counts = np.concatenate([hc[r, nmax:nmax + e - b] for r, (b, e) in enumerate(sizes)]).astype(np.int64)
    out = finalize(s, rank, host, raw_counts, sm_counts, [b for b, _ in sizes], side_counts, total_depth_mm, x_length_mm,
                   y_length_mm)
    out["sdf"] = sdf
    return out


class _MeshView:
    """Shape-compatible stand-in for engine.DeviceMesh in bench.py: this rank's slab of the stitched mesh."""

    def __init__(self, verts, faces, total_v, total_f):
        self.verts, self.faces = verts, faces
        self.total_vertices, self.total_faces = total_v, total_f
        self.n_ambiguous = 0


def gather_mesh(res: Dict, dst: int = 0, group=None):
    """Concatenate the per-rank slabs of the stitched mesh on rank `dst` (verts (V,3) f32, faces (F,3) i64)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = res["verts"].device
    sizes = torch.tensor([res["verts"].shape[0], res["faces"].shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().tolist()
    if rank == dst:
        vs, fs = [], []
        for r, (nv, nf) in enumerate(all_sizes):
            if r == rank:
                vs.append(res["verts"])
                fs.append(res["faces"])
            else:
                v = torch.empty((nv, 3), dtype=torch.float32, device=dev)
                f = torch.empty((nf, 3), dtype=torch.int64, device=dev)
                if nv:
                    dist.recv(v, r, group=group)
                if nf:
                    dist.recv(f, r, group=group)
                vs.append(v)
                fs.append(f)
        return torch.cat(vs), torch.cat(fs)
    if res["verts"].shape[0]:
        dist.send(res["verts"].contiguous(), dst, group=group)
    if res["faces"].shape[0]:
        dist.send(res["faces"].contiguous(), dst, group=group)
    return None

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


def slab_range(Z: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice range [z0, z1) of `rank`."""
    base, rem = divmod(Z, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def owned_padded_planes(Zg: int, z0: int, z1: int, pad: int = 1) -> Tuple[int, int]:
    """Planes [a, b) of the padded grid (Zg + 2*pad planes) whose lower-corner edges / cube layers a slab owns."""
    a = 0 if z0 == 0 else z0 + pad
    b = Zg + 2 * pad if z1 == Zg else z1 + pad
    return a, b


def z_map_value(plane_unpadded: float, cum: np.ndarray, adj: np.ndarray) -> np.float32:
    """float32 z coordinate the vertex transform gives to a vertex lying exactly on un-padded plane index `plane`
    (surface_extractor.py:98-113; same arithmetic as the kernel and SURVEY.md V8)."""
    z = np.float32(plane_unpadded)
    if len(cum) == 0:
        return z
    if z < 0:
        return np.float32(0)
    if z >= len(cum) - 1:
        return np.float32(cum[-1])
    lo = int(np.floor(z))
    fr = np.float32(z - np.float32(lo))
    return np.float32(cum[lo] + np.float64(fr) * adj[min(lo, len(adj) - 1)])


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
            n_ghost = ((vz == float(z_map_value(b - 1, cum, adj))) & valid).sum()
        if s.z0 > 0:
            n_lead = ((vz == float(z_map_value(a - 1, cum, adj))) & valid).sum()
    s.local = torch.cat([counts_mesh, n_ghost.reshape(1), n_lead.reshape(1), meas.view(torch.int64), bbox_t.to(torch.int64)])
    return s.local


def finalize(s: Slab, rank: int, host: torch.Tensor, raw_counts: np.ndarray, sm_counts: np.ndarray, slab_starts: List[int],
             side_counts, total_depth_mm: float, x_length_mm: float, y_length_mm: float) -> Dict:
    """Phase 5: host[r] = the vector of slab_local() of every rank; counts = global per-slice voxel counts."""
    mm_x, mm_y = x_length_mm / s.W, y_length_mm / s.H
    depths = pipeline.slice_depths(total_depth_mm, *side_counts)
    per_rank = [(int(host[r, 0]), int(host[r, 3]), int(host[r, 4])) for r in range(host.shape[0])]
    bases, consistent = stitch_offsets(per_rank)
    consistent = consistent and not bool(host[:, 2].any())   # a rank whose fast ordering failed needs the general path
    meas_all = host[:, 5:7].contiguous().view(torch.float64)
    signed_volume, area = float(meas_all[:, 0].sum()), float(meas_all[:, 1].sum())
    bbs = host[:, 7:13].numpy().copy()
    nonempty = bbs[:, 1] >= 0
    bbox = None
    if nonempty.any():
        bbs[:, 0] += np.asarray(slab_starts)
        bbs[:, 1] += np.asarray(slab_starts)
        q = bbs[nonempty]
        bbox = (int(q[:, 0].min()), int(q[:, 1].max()), int(q[:, 2].min()), int(q[:, 3].max()), int(q[:, 4].min()),
                int(q[:, 5].max()))
    v_own = per_rank[rank][0] - per_rank[rank][1]
    if s.mesh is not None:
        s.mesh.set_sizes(per_rank[rank][0], int(host[rank, 1]), 0)
        verts_own = s.mesh._verts[:v_own]
        faces_global = s.mesh._faces + bases[rank]
    else:
        verts_own = torch.empty((0, 3), dtype=torch.float32, device=s.dev)
        faces_global = torch.empty((0, 3), dtype=torch.int64, device=s.dev)
    total_v, total_f = sum(v - g for v, g, _ in per_rank), int(host[:, 1].sum())
    return {
        "verts": verts_own, "faces": faces_global, "vertex_base": bases[rank], "stitch_consistent": consistent,
        "total_vertices": total_v, "total_faces": total_f,
        "voxel_volume_mm3": pipeline.variable_depth_volume(raw_counts, mm_x, mm_y, depths),
        "processed_voxel_volume_mm3": pipeline.variable_depth_volume(sm_counts, mm_x, mm_y, depths),
        "mesh_volume_mm3": abs(signed_volume), "surface_area_mm2": area, "bbox_index": bbox,
        "active_voxels": int(raw_counts.sum()), "slice_depths": depths,
        "mesh": _MeshView(verts_own, faces_global, total_v, total_f),
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

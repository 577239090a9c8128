"""Batches of independent volumes (BASELINE configs[2]: 256 independent 256^3 phantoms over 8 GPUs).

Pure data parallelism: volume i belongs to rank i % world, every rank runs the single-GPU path on its volumes, nothing
is exchanged ("replicas").  All volumes of a batch share one FusedPlan (same shape and parameters): each
is copied into the plan's persistent input buffer and the captured CUDA graph is replayed, so a volume costs one
device-to-device copy + one graph launch + the download of its mesh."""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import numpy as np
import torch

from . import engine, pipeline


def my_items(count: int, rank: int, world: int) -> List[int]:
    """Indices of the batch items rank `rank` of `world` processes (round robin: balances a batch sorted by size)."""
    return list(range(rank, count, world))


def phantom_params(count: int, n: int, seed: int = 1234) -> Tuple[np.ndarray, np.ndarray]:
    """Radii and centres of the batch phantoms of SURVEY.md 8d: radii in [0.2, 0.45]*n, centre jitter +-0.05*n."""
    rng = np.random.default_rng(seed)
    radii = rng.uniform(0.2, 0.45, size=(count, 3)) * n
    centres = n / 2 + rng.uniform(-0.05, 0.05, size=(count, 3)) * n
    return radii, centres


def phantom_u8(n: int, radii, centre, device) -> torch.Tensor:
    """uint8 0/255 (n,n,n) ellipsoid occupancy on `device` (float64 arithmetic, same predicate as the oracle's phantom)."""
    ax = [((torch.arange(n, dtype=torch.float64, device=device) - float(c)) / float(r)) ** 2 for c, r in zip(centre, radii)]
    return ((ax[0][:, None, None] + ax[1][None, :, None] + ax[2][None, None, :]) <= 1.0).to(torch.uint8) * 255


def reconstruct_batch(stacks: Iterable, threshold: int, side_counts, total_depth_mm: float, x_length_mm: float, y_length_mm: float,
                      iterations: int = 3, close_ends: bool = True, add_padding: bool = True, rank: int = 0, world: int = 1,
                      use_graph: bool = True, keep_mesh: bool = True) -> Dict[int, Dict]:
    """stacks: sequence of (Z,H,W) uint8 stacks (CUDA tensors, or host arrays / lists of masks) of ONE shape, or a callable
    i -> stack with `len` given by stacks.count.  Returns {item index: result dict} for the items of this rank; with
    keep_mesh the canonical mesh is returned as host numpy arrays ("vertices", "faces")."""
    if callable(stacks):
        count, get = int(stacks.count), stacks
    else:
        seq = list(stacks)
        count, get = len(seq), seq.__getitem__
    dev = engine._require_cuda()
    buf = None
    out: Dict[int, Dict] = {}
    for i in my_items(count, rank, world):
        item = get(i)
        if not isinstance(item, torch.Tensor):
            item = torch.from_numpy(np.ascontiguousarray(engine._as_stack(item)))
        if buf is None:
            buf = torch.empty(tuple(item.shape), dtype=torch.uint8, device=dev)    # persistent: keeps the captured graph valid
        if tuple(item.shape) != tuple(buf.shape):
            raise ValueError("all volumes of a batch must have one shape")
        buf.copy_(item, non_blocking=True)
        r = pipeline.reconstruct_fused(buf, threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations,
                                       close_ends, add_padding, use_graph)
        mesh = r.pop("mesh")
        r["n_vertices"], r["n_faces"], r["n_ambiguous"] = int(mesh.verts.shape[0]), int(mesh.faces.shape[0]), int(mesh.n_ambiguous)
        if keep_mesh:      # the plan's output buffers are overwritten by the next item
            r["vertices"], r["faces"] = engine.download(mesh.verts), engine.download(mesh.faces)
        out[i] = r
    return out


# ----------------------------------------------------------------------------------------------------------------
# many small volumes in flight at once.  One 256^3 step is ~30 dependent kernels that each fill a fraction of the GPU:
# launch- and latency-bound (0.3 ms per volume when run one after the other).  Here every rank keeps `slots` plans,
# each with its own stream, input buffer, workspace and captured graph; volume i runs on slot i % slots, so `slots`
# graphs are in flight together and the small kernels of different volumes fill the SMs.  The host only waits for a
# slot when it needs it again.
# ----------------------------------------------------------------------------------------------------------------
class _Slot:
    __slots__ = ("plan", "stream", "event", "buf", "item")


_batch_caps: Dict = {}
_batch_slots: Dict = {}


def reconstruct_volumes(stacks, threshold: int, side_counts, total_depth_mm: float, x_length_mm: float, y_length_mm: float,
                        iterations: int = 3, close_ends: bool = True, add_padding: bool = True, keep_mesh: bool = False,
                        slots: int = 8) -> Dict[int, Dict]:
    """stacks: {item index: (Z,H,W) uint8 CUDA tensor} (or a list) of ONE shape, all on the current device.  Returns
    {item index: result dict} like reconstruct_batch.  The first call for a shape runs every volume through the staged
    path to learn the largest mesh of the batch (capacities); later calls keep `slots` volumes in flight."""
    items = stacks if isinstance(stacks, dict) else dict(enumerate(stacks))
    if not items:
        return {}
    dev = engine._require_cuda()
    first = next(iter(items.values()))
    Z, H, W = (int(v) for v in first.shape)
    key = (Z, H, W, int(threshold), tuple(side_counts), float(total_depth_mm), float(x_length_mm), float(y_length_mm),
           int(iterations), bool(close_ends), bool(add_padding), torch.cuda.current_device())
    out: Dict[int, Dict] = {}

    def finish(r: Dict, mesh) -> Dict:
        r["n_vertices"], r["n_faces"], r["n_ambiguous"] = int(mesh.verts.shape[0]), int(mesh.faces.shape[0]), int(mesh.n_ambiguous)
        if keep_mesh:
            r["vertices"], r["faces"] = engine.download(mesh.verts), engine.download(mesh.faces)
        return r

    def staged(i, item):
        r = pipeline.reconstruct(item, threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations, close_ends,
                                 add_padding)
        m = r.pop("mesh")
        caps = pipeline._caps_from(m.n_active, *m.n_raw, m.n_z)
        old = _batch_caps.get(key)
        _batch_caps[key] = caps if old is None else tuple(max(a, b) for a, b in zip(old, caps))
        out[i] = finish(r, m)

    if key not in _batch_caps:
        for i, item in items.items():
            staged(i, item)
        return out
    caps = pipeline._tuned_caps(("batch",) + key, _batch_caps[key])
    state = _batch_slots.get(key)
    if state is None or state[0] != caps or len(state[1]) != slots:
        ring = []
        for _ in range(slots):
            s = _Slot()
            s.plan = pipeline.FusedPlan((Z, H, W), threshold, side_counts, total_depth_mm, x_length_mm, y_length_mm, iterations,
                                        close_ends, add_padding, caps, dev)
            s.stream, s.event, s.item = torch.cuda.Stream(device=dev), torch.cuda.Event(), None
            s.buf = torch.empty((Z, H, W), dtype=torch.uint8, device=dev)
            with torch.cuda.stream(s.stream):
                s.plan.capture(s.buf)
            s.stream.synchronize()
            ring.append(s)
        state = _batch_slots[key] = (caps, ring)
    ring = state[1]
    R = pipeline
    redo = []

    def harvest(s):
        if s.item is None:
            return
        s.event.synchronize()
        r = s.plan.res_np
        h = r[:R.R_COUNTS].tolist()
        if h[R.R_OVERFLOW] or h[R.R_UNVERIFIED] or h[R.R_NT] == 0:
            redo.append(s.item)          # larger than anything seen so far (or an unverifiable fast ordering): staged path
        else:
            res = R._result_from_block(s.plan, r, h)
            mesh = res.pop("mesh")
            with torch.cuda.stream(s.stream):
                out[s.item] = finish(res, mesh)
        s.item = None

    main = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(main)
    for k, (i, item) in enumerate(items.items()):
        s = ring[k % slots]
        harvest(s)
        with torch.cuda.stream(s.stream):
            s.stream.wait_event(ready)
            s.buf.copy_(item, non_blocking=True)
            s.plan.launch(s.buf, True)
            s.event.record(s.stream)
        s.item = i
    for s in ring:
        harvest(s)
        main.wait_stream(s.stream)
    for i in redo:
        staged(i, items[i])
    return out

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

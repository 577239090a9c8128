"""Whole-job pipeline on device-resident inputs: uint8 mask stack -> voxel grid -> smoothing -> mesh + volumes.

This is the order tomography_3d_reconstruction.py runs the hot path in (create_voxel_data :88-100, calculate_volume
:102-118, smooth + extract + mesh volume :120-140, surface area :207-223, analyze :225-229), run once per step
instead of the orchestrator's 5x smooth / 4x extract on identical inputs.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import engine


def slice_depths(total_depth_mm: float, side_0: int, side_1: int, side_2: int) -> np.ndarray:
    """voxel_processor.py:129-163 (host scalars)."""
    total = side_0 + side_1 + side_2
    if side_1 == 0 or total == 0:
        return np.array([]) if total == 0 else np.full(total, total_depth_mm / total)
    d1 = total_depth_mm / side_1
    d02 = 2 * d1
    d0 = d02 / side_0 if side_0 > 0 else 0
    d2 = d02 / side_2 if side_2 > 0 else 0
    return np.array([d0] * side_0 + [d1] * side_1 + [d2] * side_2)


def variable_depth_volume(counts: np.ndarray, mm_x: float, mm_y: float, depths: np.ndarray) -> float:
    """volume_calculator.py:23-35 on exact per-slice counts, same float64 order."""
    if len(depths) == 0:
        return 0.0
    n = min(len(counts), len(depths))
    if n == 0:
        return 0.0
    # the reference's loop `total += count[z] * (mm_x * mm_y * depth[z])`, vectorised without changing a bit:
    # the products are the same float64 operations and np.cumsum accumulates strictly left to right
    prod = np.asarray(counts[:n], dtype=np.int64).astype(np.float64) * ((mm_x * mm_y) * np.asarray(depths[:n], dtype=np.float64))
    return float(np.cumsum(prod)[-1])


def reconstruct(masks_u8: torch.Tensor, threshold: int, side_counts, total_depth_mm: float, x_length_mm: float,
                y_length_mm: float, iterations: int = 3, close_ends: bool = True, add_padding: bool = True,
                mark: Optional[Callable[[str], None]] = None) -> Dict:
    """masks_u8: CUDA uint8 (Z,H,W).  Returns the device mesh and the host scalars of analyze_object_properties.

    `mark(name)` is called after each stage has been enqueued (bench.py records CUDA events there)."""
    mark = mark or (lambda _n: None)
    Z, H, W = (int(s) for s in masks_u8.shape)
    mm_x, mm_y = x_length_mm / W, y_length_mm / H
    depths = slice_depths(total_depth_mm, *side_counts)

    dv = engine.pack_and_close(masks_u8, threshold, close_ends)
    mark("pack_close")
    # bounding box of the raw grid: off the critical path
    main = torch.cuda.current_stream()
    side = engine.side_stream(masks_u8.device)
    ready = torch.cuda.Event()
    ready.record(main)
    with torch.cuda.stream(side):
        side.wait_event(ready)
        bbox_t = dv.bbox_tensor()
        bbox_done = torch.cuda.Event()
        bbox_done.record(side)
    sm = engine.smooth(dv, iterations, True)
    mark("smooth")
    mesh = engine.extract_surface(sm, depths, mm_y, mm_x, True, add_padding, canonical="async", mark=mark)
    # measures on the emitted (pre-canonical) mesh: merging exact duplicates and dropping zero-area faces changes
    # neither the signed volume nor the area, and it removes a dependency on the canonical sizes
    raw_verts, raw_faces = mesh.raw
    meas = engine.mesh_measure_async(raw_verts, raw_faces)
    mark("measure")
    main.wait_event(bbox_done)
    # one device->host copy for every scalar result of the step
    Zc = dv.Z
    packed = torch.cat([mesh.counts_dev, meas.view(torch.int64), dv.counts_tensor(), sm.counts_tensor(),
                        bbox_t.to(torch.int64)]).cpu()
    mark("stats")
    mesh.set_sizes(int(packed[0]), int(packed[1]), int(packed[2]))
    signed_volume, area = (float(x) for x in packed[3:5].view(torch.float64).tolist())
    mesh._measures = (signed_volume, area)
    raw_counts = packed[5:5 + Zc].numpy().astype(np.int64)
    sm_counts = packed[5 + Zc:5 + 2 * Zc].numpy().astype(np.int64)
    bb = tuple(int(x) for x in packed[5 + 2 * Zc:].tolist())
    dv.set_host_stats(raw_counts, bb)
    sm.set_host_stats(sm_counts, None)
    bbox = dv.bbox()
    return {
        "mesh": mesh,
        "voxel_volume_mm3": variable_depth_volume(raw_counts, mm_x, mm_y, depths),
        "processed_voxel_volume_mm3": variable_depth_volume(sm_counts, mm_x, mm_y, depths),
        "mesh_volume_mm3": abs(signed_volume),
        "surface_area_mm2": area,
        "bbox_index": bbox,
        "active_voxels": int(raw_counts.sum()),
        "slice_depths": depths,
    }

#!/usr/bin/env python3
"""Timings of the BASELINE configurations other than the headline one, on ONE GPU (what a single rank of the multi-GPU
configurations does): one JSON line each, CUDA-event timed, synthetic phantoms.

    python tools/extra_bench.py [c0] [c2] [c3slab] [c4slab] > profiles/rNN_configs.jsonl
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tomography_3d_reconstructor_b200 import batch, pipeline  # noqa: E402

PHYS = (bench.PHYS["total_depth_mm"], bench.PHYS["x_length_mm"], bench.PHYS["y_length_mm"])


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def occupancy(name, Z, H, W, sides, steps):
    dev = torch.device("cuda", 0)
    masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
    fn = lambda: pipeline.reconstruct_fused(masks, bench.THRESHOLD, sides, *PHYS)
    fn()
    ms, out = timed(fn, steps)
    m = out["mesh"]
    print(json.dumps({"config": name, "shape": [Z, H, W], "path": "occupancy, fused + CUDA graph", "ms_per_step": ms,
                      "Gvoxels/s": Z * H * W / ms / 1e6, "vertices": int(m.verts.shape[0]), "faces": int(m.faces.shape[0]),
                      "mesh_volume_mm3": out["mesh_volume_mm3"], "voxel_volume_mm3": out["voxel_volume_mm3"]}), flush=True)
    del masks
    pipeline._plans.clear()
    torch.cuda.empty_cache()


def main():
    which = set(sys.argv[1:]) or {"c0", "c2", "c3slab", "c4slab"}
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    if "c0" in which:
        occupancy("C0: 104 x 512 x 512 (Section_0/1/2 = 20+64+20)", 104, 512, 512, (20, 64, 20), 20)
    if "c2" in which:
        n, count = 256, 32                       # one GPU's share of the 256-phantom batch over 8 GPUs
        radii, centres = batch.phantom_params(256, n)
        stacks = [batch.phantom_u8(n, radii[i], centres[i], dev) for i in range(count)]
        sides = bench.side_counts(n)
        fn = lambda: batch.reconstruct_batch(stacks, bench.THRESHOLD, sides, *PHYS, keep_mesh=False)
        fn()
        ms, out = timed(fn, 3, 1)
        print(json.dumps({"config": "C2: batch of independent 256^3 phantoms (one GPU's 32 of 256)", "shape": [count, n, n, n],
                          "path": "batch.reconstruct_batch: shared plan + CUDA graph, one D2D copy + one launch per volume",
                          "ms_per_batch": ms, "ms_per_volume": ms / count, "Gvoxels/s": count * n ** 3 / ms / 1e6,
                          "faces_total": int(sum(o["n_faces"] for o in out.values()))}), flush=True)
        del stacks
        pipeline._plans.clear()
        torch.cuda.empty_cache()
    if "c3slab" in which:
        Z, H, W = 256, 2048, 2048                # one GPU's z-slab of the 2048^3 stack over 8 GPUs
        masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
        sides = bench.side_counts(Z)
        fn = lambda: pipeline.reconstruct_sdf(masks, bench.THRESHOLD, sides, *PHYS)
        ms, out = timed(fn, 2, 1)
        m = out["mesh"]
        print(json.dumps({"config": "C3 slab: 256 x 2048 x 2048 with exact EDT/SDF + marching cubes on the distance field",
                          "shape": [Z, H, W], "path": "pipeline.reconstruct_sdf (staged)", "ms_per_step": ms,
                          "Gvoxels/s": Z * H * W / ms / 1e6, "vertices": int(m.verts.shape[0]), "faces": int(m.faces.shape[0]),
                          "mesh_volume_mm3": out["mesh_volume_mm3"], "voxel_volume_mm3": out["processed_voxel_volume_mm3"]}), flush=True)
        del masks, out, m
        torch.cuda.empty_cache()
    if "c4slab" in which:
        occupancy("C4 slab: 512 x 4096 x 4096 (one GPU's z-slab of the 4096^3 stack over 8 GPUs)", 512, 4096, 4096,
                  bench.side_counts(512), 5)


if __name__ == "__main__":
    main()

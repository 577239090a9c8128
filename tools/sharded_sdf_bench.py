#!/usr/bin/env python3
"""Timing of the z-slab sharded SDF pipeline (BASELINE configs[3]: exact EDT/SDF + marching cubes on the distance field),
run under torchrun like bench.py.  Per-GPU slab shape from argv (default 256,2048,2048 = 2048^3 over 8 GPUs)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/t3d_nccl.%h.%p.log")
    dist.init_process_group("nccl", device_id=dev)
    from tomography_3d_reconstructor_b200 import sharded
    Z, H, W = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,2048").split(","))
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    Zg = Z * world
    sides = bench.side_counts(Zg)
    phys = (bench.PHYS["total_depth_mm"], bench.PHYS["x_length_mm"], bench.PHYS["y_length_mm"])
    z0, z1 = sharded.slab_range(Zg, rank, world)
    masks = bench.make_phantom_u8(Zg, H, W, z0, z1, dev)
    out = sharded.reconstruct_sdf(masks, Zg, z0, bench.THRESHOLD, sides, *phys)
    del out
    torch.cuda.empty_cache()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = sharded.reconstruct_sdf(masks, Zg, z0, bench.THRESHOLD, sides, *phys)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"config": "C3: %d x %d x %d stack, exact EDT/SDF + marching cubes, z-slab sharded" % (Zg, H, W),
                          "n_gpus": world, "path": "sharded.reconstruct_sdf (staged; EDT all-to-all transpose over NCCL)",
                          "ms_per_step": ms, "Gvoxels/s": Zg * H * W / ms / 1e6, "steps": steps,
                          "vertices": out["total_vertices"], "faces": out["total_faces"],
                          "mesh_volume_mm3": out["mesh_volume_mm3"], "voxel_volume_mm3": out["processed_voxel_volume_mm3"],
                          "stitch_consistent": bool(out["stitch_consistent"])}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

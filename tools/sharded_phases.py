#!/usr/bin/env python3
"""Per-rank phase times of the fused sharded step (eager, CUDA events): pack | halo exchange | t3d_reconstruct_slab |
all-gather | stitch | D2H+sync+assemble (host).  Run under torchrun like bench.py."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    from tomography_3d_reconstructor_b200 import sharded
    Z, H, W = 512, 1024, 1024
    Zg = Z * world
    sides = bench.side_counts(Zg)
    phys = (bench.PHYS["total_depth_mm"], bench.PHYS["x_length_mm"], bench.PHYS["y_length_mm"])
    z0, z1 = sharded.slab_range(Zg, rank, world)
    masks = bench.make_phantom_u8(Zg, H, W, z0, z1, dev)
    for _ in range(3):
        out = sharded.reconstruct_fused(masks, Zg, z0, bench.THRESHOLD, sides, *phys, use_graph=False)
    plan = next(iter(sharded._slab_plans.values()))
    names = ["pack", "halo", "compute", "gather", "stitch"]
    acc = np.zeros(len(names) + 2)
    steps = 10
    for _ in range(steps):
        dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        t0 = time.perf_counter()
        ev[0].record()
        plan.pack(masks); ev[1].record()
        sharded.exchange_halos(plan.ext, plan.hl, plan.n, plan.hh, rank, world, None); ev[2].record()
        plan.compute(); ev[3].record()
        dist.all_gather_into_tensor(plan.gathered, plan.res); ev[4].record()
        plan.stitch(); ev[5].record()
        t1 = time.perf_counter()
        plan.host.copy_(plan.gathered, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        t2 = time.perf_counter()
        sharded.assemble(plan, plan.host.numpy())
        t3 = time.perf_counter()
        acc[:len(names)] += [ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))]
        acc[-2] += 1e3 * (t1 - t0)
        acc[-1] += 1e3 * (t3 - t2)
    acc /= steps
    mesh = int(plan.host[rank, 16])
    allr = [None] * world
    dist.all_gather_object(allr, (rank, acc.tolist(), mesh))
    if rank == 0:
        print("rank  " + "  ".join("%8s" % n for n in names) + "   enqueue_host  assemble_host   raw_verts")
        for r, a, m in sorted(allr):
            print("%4d  " % r + "  ".join("%8.3f" % v for v in a[:len(names)]) + "   %10.3f  %12.3f  %10d" % (a[-2], a[-1], m))
    sharded._slab_plans.clear()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

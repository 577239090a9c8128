#!/usr/bin/env python3
"""Time the phases of the sharded step of ONE rank (world = 1: no exchange) with CUDA events: pack() and compute() of the
FusedSlabPlan, in the pre-filled mode (default) or the plain mode (T3D_NO_SLAB_PACK_GAP=1).
    python tools/slab_mode_probe.py Z H W"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tomography_3d_reconstructor_b200 import sharded  # noqa: E402

Z, H, W = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
import torch.distributed as dist  # noqa: E402
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1)
masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
sides = bench.side_counts(Z)
phys = (bench.PHYS["total_depth_mm"], bench.PHYS["x_length_mm"], bench.PHYS["y_length_mm"])
for _ in range(3):
    out = sharded.reconstruct_fused(masks, Z, 0, bench.THRESHOLD, sides, *phys, use_graph=False)
plan = next(iter(sharded._slab_plans.values()))
plan = plan["plan"] if isinstance(plan, dict) else plan
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for rep in range(4):
    torch.cuda.synchronize()
    ev[0].record()
    plan.pack(masks)
    ev[1].record()
    plan.compute()
    ev[2].record()
    torch.cuda.synchronize()
    print("rep %d: pre_active=%s pack %.1f us  compute %.1f us  total %.1f us" % (rep, plan.pre_active, 1e3 * ev[0].elapsed_time(ev[1]),
          1e3 * ev[1].elapsed_time(ev[2]), 1e3 * ev[0].elapsed_time(ev[2])), flush=True)
dist.destroy_process_group()

#!/usr/bin/env python3
"""Pure device time of the captured C1 step: K graph replays back to back (no host work in between) vs the full step."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tomography_3d_reconstructor_b200 import pipeline
Z, H, W = 512, 1024, 1024
dev = torch.device("cuda", 0)
masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
sides = bench.side_counts(Z)
args = (masks, bench.THRESHOLD, sides, bench.PHYS["total_depth_mm"], bench.PHYS["x_length_mm"], bench.PHYS["y_length_mm"])
for _ in range(4):
    pipeline.reconstruct_fused(*args)
plan = next(iter(pipeline._plans.values()))
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(K):
    plan.graph.replay()
e1.record()
torch.cuda.synchronize()
print("graph replays back to back: %.1f us per step" % (1e3 * e0.elapsed_time(e1) / K))
e0.record()
for _ in range(K):
    pipeline.reconstruct_fused(*args)
e1.record()
torch.cuda.synchronize()
print("full reconstruct_fused    : %.1f us per step" % (1e3 * e0.elapsed_time(e1) / K))
t0 = time.perf_counter()
for _ in range(K):
    r = plan.run(masks, True)
t1 = time.perf_counter()
print("plan.run (replay+D2H+sync): %.1f us per step (wall)" % (1e6 * (t1 - t0) / K))

#!/usr/bin/env python3
"""Host-side vs device-side time per stage of pipeline.reconstruct (is the step launch-bound?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from tomography_3d_reconstructor_b200 import pipeline

Z, H, W = 512, 1024, 1024
dev = torch.device("cuda", 0)
masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
sides = bench.side_counts(Z)
for _ in range(5):
    pipeline.reconstruct(masks, 200, sides, 6.0, 143.1, 95.03)
torch.cuda.synchronize()
acc = {}
N = 20
for _ in range(N):
    marks = []
    def mark(name):
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))
    torch.cuda.synchronize()
    mark("start")
    pipeline.reconstruct(masks, 200, sides, 6.0, 143.1, 95.03, mark=mark)
    mark("end")
    torch.cuda.synchronize()
    for (n0, e0, t0), (n1, e1, t1) in zip(marks[:-1], marks[1:]):
        a = acc.setdefault(n1, [0.0, 0.0]); a[0] += e0.elapsed_time(e1); a[1] += (t1 - t0) * 1e3
print("%-14s %10s %10s" % ("stage", "device_ms", "host_ms"))
td = th = 0
for k, (d, h) in acc.items():
    print("%-14s %10.3f %10.3f" % (k, d / N, h / N)); td += d / N; th += h / N
print("%-14s %10.3f %10.3f" % ("total", td, th))

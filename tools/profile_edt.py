#!/usr/bin/env python3
"""Signed distance of the smoothed C1-shaped phantom between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tomography_3d_reconstructor_b200 import edt, engine  # noqa: E402

Z, H, W = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "512,1024,1024").split(","))
dev = torch.device("cuda", 0)
masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
sm = engine.smooth(engine.pack_and_close(masks, bench.THRESHOLD, True), 3, True)
del masks
samp = (0.0117, 0.0928, 0.1397)
for _ in range(2):
    sdf = edt.signed_distance(sm, samp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
sdf = edt.signed_distance(sm, samp)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("sdf %dx%dx%d: %.2f ms, %.1f Gvoxels/s  min %.4f max %.4f" % (Z, H, W, e0.elapsed_time(e1), Z * H * W / e0.elapsed_time(e1) / 1e6,
                                                                   float(sdf.min()), float(sdf.max())))

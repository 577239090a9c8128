#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 T3D_RAISE=1 python - <<'PY' > gpurun_out/r2p_dev1.log 2>&1
import sys, faulthandler
sys.path.insert(0, ".")
faulthandler.enable()
import numpy as np, torch
from oracle import cpu_ref as oracle
from tomography_3d_reconstructor_b200 import pipeline, engine
Z, H, W = 24, 64, 96
u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
sides = (3, 18, 3)
for d in (0, 1):
    torch.cuda.set_device(d)
    masks = torch.from_numpy(u8).to("cuda:%d" % d)
    for rep in range(3):
        print("device", d, "rep", rep, flush=True)
        if rep == 0:
            out = pipeline.reconstruct(masks, 200, sides, 6.0, 143.1, 95.03)
        else:
            out = pipeline.reconstruct_fused(masks, 200, sides, 6.0, 143.1, 95.03, use_graph=(rep == 2))
        torch.cuda.synchronize()
        print("   ok", out["mesh"].verts.shape, flush=True)
PY
tail -40 gpurun_out/r2p_dev1.log | cut -c1-220

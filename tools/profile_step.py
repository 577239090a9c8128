#!/usr/bin/env python3
"""One fused (non-graph) C1 step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`:

    python tools/profile_step.py                      # plain run first (must exit 0)
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof \
        python tools/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tomography_3d_reconstructor_b200 import pipeline  # noqa: E402


def main():
    Z, H, W = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "512,1024,1024").split(","))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
    sides = bench.side_counts(Z)
    args = (masks, bench.THRESHOLD, sides, bench.PHYS["total_depth_mm"], bench.PHYS["x_length_mm"], bench.PHYS["y_length_mm"])
    for _ in range(3):   # staged (learns the sizes), then fused
        res = pipeline.reconstruct_fused(*args, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    res = pipeline.reconstruct_fused(*args, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    m = res["mesh"]
    print("V=%d F=%d volume=%.6f" % (m.verts.shape[0], m.faces.shape[0], res["mesh_volume_mm3"]))


if __name__ == "__main__":
    main()

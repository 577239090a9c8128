#!/usr/bin/env python3
"""Size-independent property checks at shapes the CPU oracle cannot reach (64-bit indexing, wide rows)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from tomography_3d_reconstructor_b200 import pipeline

dev = torch.device("cuda", 0)
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(64, 4096, 4096), (1024, 2048, 2048)]
for (Z, H, W) in shapes:
    masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
    sides = bench.side_counts(Z)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = pipeline.reconstruct(masks, 200, sides, 6.0, 143.1, 95.03)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pipeline._plans.clear(); pipeline._hints.clear()
    pipeline.reconstruct_fused(masks, 200, sides, 6.0, 143.1, 95.03)          # learns sizes (staged)
    b = pipeline.reconstruct_fused(masks, 200, sides, 6.0, 143.1, 95.03)      # fused + graph
    torch.cuda.synchronize(); t2 = time.perf_counter()
    b = pipeline.reconstruct_fused(masks, 200, sides, 6.0, 143.1, 95.03)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    va, fa = a["mesh"].verts, a["mesh"].faces
    vb, fb = b["mesh"].verts, b["mesh"].faces
    same = bool(torch.equal(va, vb) and torch.equal(fa, fb)) and a["voxel_volume_mm3"] == b["voxel_volume_mm3"] \
        and a["processed_voxel_volume_mm3"] == b["processed_voxel_volume_mm3"] and a["bbox_index"] == b["bbox_index"]
    V, F = int(va.shape[0]), int(fa.shape[0])
    # closed oriented 2-manifold: every directed edge once, its twin exists; Euler characteristic 2
    e = torch.cat([fa[:, [0, 1]], fa[:, [1, 2]], fa[:, [2, 0]]])
    key = e[:, 0] * (V + 1) + e[:, 1]
    rev = e[:, 1] * (V + 1) + e[:, 0]
    uk = torch.unique(key)
    closed = bool(uk.numel() == key.numel() and torch.equal(torch.sort(key)[0], torch.sort(rev)[0]))
    euler = V - key.numel() // 2 + F
    d = (va[1:].double() - va[:-1].double())
    lex = bool(((d[:, 0] > 0) | ((d[:, 0] == 0) & ((d[:, 1] > 0) | ((d[:, 1] == 0) & (d[:, 2] > 0))))).all())
    analytic_vox = 4 / 3 * np.pi * (0.42 * Z) * (0.33 * H) * (0.45 * W)
    print("%dx%dx%d  V=%d F=%d  staged %.1f ms  fused %.2f ms  fused==staged %s  closed %s  euler %d  sorted %s  active/analytic %.5f  amb %d" % (
        Z, H, W, V, F, (t1 - t0) * 1e3, (t3 - t2) * 1e3, same, closed, euler, lex, a["active_voxels"] / analytic_vox, a["mesh"].n_ambiguous), flush=True)
    assert same and closed and euler == 2 and lex
    del masks, a, b, va, vb, fa, fb, e, key, rev, uk, d
    pipeline._plans.clear(); pipeline._hints.clear()
    torch.cuda.empty_cache()
print("big_check ok")

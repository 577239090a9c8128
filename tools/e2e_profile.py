#!/usr/bin/env python3
"""Wall time of each reference-facing API call of one e2e step (host buffers in pinned memory)."""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor, VolumeCalculator

Z, H, W = 512, 1024, 1024
dev = torch.device("cuda", 0)
masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
host_bool = torch.empty((Z, H, W), dtype=torch.bool, pin_memory=True)
host_bool.copy_(masks >= 200)
torch.cuda.synchronize()
hb = host_bool.numpy()
mask_list = [hb[z] for z in range(Z)]
sides = bench.side_counts(Z)
mm_x, mm_y = 143.1 / W, 95.03 / H
acc = {}
def T(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
with contextlib.redirect_stdout(io.StringIO()):
    for it in range(5):
        vp, se, vc = VoxelProcessor(), SurfaceExtractor(), VolumeCalculator()
        vox = T("create_voxel_data", lambda: vp.create_voxel_data(mask_list, True, *sides))
        depths = vp.calculate_slice_depths(6.0)
        sm = T("smooth_voxel_data", lambda: vp.smooth_voxel_data(vox, 3, True))
        pv = T("voxel_volume(sm)", lambda: vc.calculate_voxel_volume_variable_depth(sm, mm_x, mm_y, depths))
        v, f = T("extract_manifold_surface", lambda: se.extract_manifold_surface(sm, depths, mm_y, mm_x, True, True, True))
        mv = T("calculate_mesh_volume", lambda: se.calculate_mesh_volume(v, f))
        ar = T("calculate_surface_area", lambda: se.calculate_surface_area(v, f))
        T("analyze_object_properties", lambda: vc.analyze_object_properties(vox, pv, mv, ar, mm_x, mm_y, depths, 143.1, 95.03, 6.0))
for k, v in acc.items():
    print("%-28s %8.2f ms (first %.2f)" % (k, float(np.mean(v[2:])), v[0]))
print("total %.2f ms" % sum(float(np.mean(v[2:])) for v in acc.values()))
# pieces of create_voxel_data
from tomography_3d_reconstructor_b200 import engine
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); st = engine._as_stack(mask_list); t1 = time.perf_counter()
    d = engine.upload_u8(st); torch.cuda.synchronize(); t2 = time.perf_counter()
    dv = engine.pack_and_close(d, 1, True); torch.cuda.synchronize(); t3 = time.perf_counter()
    hst = dv.to_host(); t4 = time.perf_counter()
    print("as_stack %.2f upload %.2f pack_close %.2f to_host %.2f" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))

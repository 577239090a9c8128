#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2m_classes.log 2>&1
import sys, time, io, contextlib
sys.path.insert(0, ".")
import torch, numpy as np
import bench
from tomography_3d_reconstructor_b200 import VoxelProcessor, SurfaceExtractor, VolumeCalculator, engine
dev = torch.device("cuda", 0)
Z, H, W = 512, 1024, 1024
masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
host_bool = torch.empty((Z, H, W), dtype=torch.bool, pin_memory=True)
host_bool.copy_(masks >= 200); torch.cuda.synchronize()
hb = host_bool.numpy()
mask_list = [hb[z] for z in range(Z)]
src = engine.mask_source(mask_list)
print("mask_source:", type(src[0]), src[1:], "base is hb:", mask_list[3].base is hb, type(mask_list[3].base))
sides = bench.side_counts(Z); mm_x, mm_y = 143.1 / W, 95.03 / H
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    with contextlib.redirect_stdout(io.StringIO()):
        vp, se, vc = VoxelProcessor(), SurfaceExtractor(), VolumeCalculator()
        t0 = T(); vox = vp.create_voxel_data(mask_list, True, *sides)
        t1 = T(); depths = vp.calculate_slice_depths(6.0); sm = vp.smooth_voxel_data(vox, 3, True)
        t2 = T(); processed = vc.calculate_voxel_volume_variable_depth(sm, mm_x, mm_y, depths)
        t3 = T(); v, f = se.extract_manifold_surface(sm, depths, mm_y, mm_x, True, True, True)
        t4 = T(); mv = se.calculate_mesh_volume(v, f); ar = se.calculate_surface_area(v, f)
        t5 = T(); props = vc.analyze_object_properties(vox, processed, mv, ar, mm_x, mm_y, depths, 143.1, 95.03, 6.0)
        t6 = T()
    print("rep %d: create %.1f smooth %.1f volume %.1f extract %.1f measures %.1f analyze %.1f total %.1f ms" %
          (rep, *(1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6), (t0, t6)))))
PY
cat gpurun_out/r2m_classes.log
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2m_tests.log
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2m_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2m_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2m_launches.csv 8 | grep "k_mc_words\|k_mc_emit"
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r2m_c1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2m_c1.json').read().strip().splitlines()[-1]); print('c1', d['value'], d['ms_per_step'])"

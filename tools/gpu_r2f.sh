#!/bin/bash
# round-2 GPU call F: SDF kernel iteration (tests + timing in both tile-loader modes + per-kernel times)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edt.py tests/test_gpu_pipeline.py -m gpu -q -x -k "edt or sdf" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2f_tests.log
T3D_SDF_NO_BULK=1 timeout 900 python -m pytest tests/test_gpu_edt.py -m gpu -q -x > gpurun_out/r2f_tests_nobulk.log 2>&1; echo "nobulk tests rc=$?"
tail -3 gpurun_out/r2f_tests_nobulk.log
cat > /tmp/sdf_time.py <<'PY'
import torch, time, sys
sys.path.insert(0, ".")
import bench
from tomography_3d_reconstructor_b200 import engine, edt
dev = torch.device("cuda", 0)
shapes = ((512, 1024, 1024), (256, 2048, 2048)) if len(sys.argv) < 2 else ((512, 1024, 1024),)
for (Z, H, W) in shapes:
    masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
    dv = engine.smooth(engine.pack_and_close(masks, 200, True), 3, True)
    del masks
    for name, fn in (("one sweep", edt.signed_distance),):
        samp = (6.0 / Z, 95.03 / H, 143.1 / W)
        out = fn(dv, samp); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = fn(dv, samp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("%dx%dx%d %-15s %.2f ms  %.1f Gvox/s  %.0f GB/s at 17.5 B/voxel" % (Z, H, W, name, ms, Z * H * W / ms / 1e6, 17.5 * Z * H * W / ms / 1e6), flush=True)
        del out
    del dv
    torch.cuda.empty_cache()
PY
python /tmp/sdf_time.py > gpurun_out/r2f_sdf.log 2>&1; cat gpurun_out/r2f_sdf.log
T3D_SDF_NO_BULK=1 python /tmp/sdf_time.py > gpurun_out/r2f_sdf_nobulk.log 2>&1; echo "--- no bulk"; cat gpurun_out/r2f_sdf_nobulk.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sdf -c 12 --csv --log-file gpurun_out/r2f_sdf_ncu.csv python /tmp/sdf_time.py c1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/r2f_sdf_ncu.csv") if not l.startswith("=="))]
by={}
for r in rows:
    by.setdefault((r["ID"], r["Kernel Name"][:40]), {})[r["Metric Name"]]=r["Metric Value"]+" "+r["Metric Unit"]
for k,v in list(by.items())[:9]:
    print(k, v)
PY

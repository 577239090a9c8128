#!/bin/bash
# round-2 GPU call E: EDT/SDF v3 (merged sweeps + bulk-async tiles), generator, orchestrator
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edt.py tests/test_gpu_generator.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -15 gpurun_out/r2e_tests.log
T3D_SDF_NO_BULK=1 timeout 900 python -m pytest tests/test_gpu_edt.py -m gpu -q -x > gpurun_out/r2e_tests_nobulk.log 2>&1; echo "nobulk tests rc=$?"
tail -3 gpurun_out/r2e_tests_nobulk.log
if [ -d oracle/_ref/reference_checkout ]; then
  python tools/run_orchestrator.py --reference oracle/_ref/reference_checkout --log gpurun_out/r2e_orchestrator.log > gpurun_out/r2e_orch.out 2>&1; echo "orchestrator rc=$?"
  tail -12 gpurun_out/r2e_orch.out
fi
cat > /tmp/sdf_time.py <<'PY'
import torch, time, sys
sys.path.insert(0, ".")
import bench
from tomography_3d_reconstructor_b200 import engine, edt
dev = torch.device("cuda", 0)
for (Z, H, W) in ((512, 1024, 1024), (256, 2048, 2048)):
    masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
    dv = engine.smooth(engine.pack_and_close(masks, 200, True), 3, True)
    del masks
    for name, fn in (("one sweep", edt.signed_distance), ("two transforms", edt.signed_distance_two_transforms)):
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
python /tmp/sdf_time.py > gpurun_out/r2e_sdf.log 2>&1; cat gpurun_out/r2e_sdf.log
T3D_SDF_NO_BULK=1 python /tmp/sdf_time.py > gpurun_out/r2e_sdf_nobulk.log 2>&1; echo "--- no bulk"; cat gpurun_out/r2e_sdf_nobulk.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2e_sdf_launches.csv python /tmp/sdf_time.py > /dev/null 2>&1
grep -c k_sdf gpurun_out/r2e_sdf_launches.csv

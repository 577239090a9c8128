#!/bin/bash
# round-2 GPU call D: full GPU suite, orchestrator harness, benches (C1, C3 SDF), SDF timing at C1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -15 gpurun_out/r2d_tests.log
if [ -d oracle/_ref/reference_checkout ]; then
  python tools/run_orchestrator.py --reference oracle/_ref/reference_checkout --log gpurun_out/r2d_orchestrator.log > gpurun_out/r2d_orch.out 2>&1; echo "orchestrator rc=$?"
  tail -12 gpurun_out/r2d_orch.out
fi
python bench.py --no-cpu > gpurun_out/r2d_c1.json 2> gpurun_out/r2d_c1.err; echo "c1 rc=$?"
python bench.py --config C3 --steps 3 > gpurun_out/r2d_c3.json 2> gpurun_out/r2d_c3.err; echo "c3 rc=$?"
python - <<'PY' > gpurun_out/r2d_sdf.log 2>&1
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
        print("%dx%dx%d %-15s %.2f ms  %.1f Gvox/s  %.0f GB/s at 17.5 B/voxel" % (Z, H, W, name, ms, Z * H * W / ms / 1e6, 17.5 * Z * H * W / ms / 1e6))
        del out
    del dv
    torch.cuda.empty_cache()
PY
cat gpurun_out/r2d_sdf.log
for f in c1 c3; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("value"), d["stages_ms"])
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2d_$f.err").read()[-1500:])
PY
done

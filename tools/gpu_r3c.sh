#!/bin/bash
# batched loads in the per-layer radix sort; eager-launch stage timeline of the fused step (T3D_STAGE_EVENTS)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_surface.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r3c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3c_tests.log
tail -4 gpurun_out/r3c_tests.log
line() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1200:])
PY
}
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r3c_c1.json 2> gpurun_out/r3c_c1.err; line r3c_c1
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r3c_c4.json 2> gpurun_out/r3c_c4.err; line r3c_c4
T3D_STAGE_EVENTS=1 python bench.py --steps 5 --no-cpu --no-e2e --no-graph --no-check > gpurun_out/r3c_c1_eager.json 2> gpurun_out/r3c_c1_eager.err; line r3c_c1_eager; grep "t3d stages" gpurun_out/r3c_c1_eager.err | tail -2
T3D_STAGE_EVENTS=1 python bench.py --config C4 --steps 3 --no-cpu --no-e2e --no-graph --no-check > gpurun_out/r3c_c4_eager.json 2> gpurun_out/r3c_c4_eager.err; line r3c_c4_eager; grep "t3d stages" gpurun_out/r3c_c4_eager.err | tail -2

#!/bin/bash
# k_morph4 v2: parity subset + C1/C4 lines + ncu --set full of the kernel alone
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2x_tests.log
tail -4 gpurun_out/r2x_tests.log
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
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r2x_c1.json 2> gpurun_out/r2x_c1.err; line r2x_c1
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2x_c4.json 2> gpurun_out/r2x_c4.err; line r2x_c4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_morph4 --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2x_morph4_c4 python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r2x_ncu.log 2>&1; echo "ncu rc=$?"

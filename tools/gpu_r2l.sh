#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_edt.py -m gpu -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
python bench.py > gpurun_out/r2l_c1.json 2> gpurun_out/r2l_c1.err; echo "c1 rc=$?"
python bench.py --config C0 > gpurun_out/r2l_c0.json 2> gpurun_out/r2l_c0.err; echo "c0 rc=$?"
for f in c1 c0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "classes", (d.get("e2e_classes") or {}).get("ms_per_step"), "bits", (d.get("e2e_bits") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2l_$f.err").read()[-1500:])
PY
done

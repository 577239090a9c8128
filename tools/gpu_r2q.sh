#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2q_tests.log
tail -6 gpurun_out/r2q_tests.log
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2q_c4.json 2> gpurun_out/r2q_c4.err
python bench.py --steps 20 --no-cpu > gpurun_out/r2q_c1.json 2> gpurun_out/r2q_c1.err
for f in c4 c1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2q_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "classes", (d.get("e2e_classes") or {}).get("ms_per_step"))
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2q_$f.err").read()[-1500:])
PY
done

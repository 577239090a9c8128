#!/bin/bash
# round-2 GPU call B: GPU test suite (MC33 resolution, in-flight batches, full-size parity) + driver-style bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
python bench.py > gpurun_out/r2b_c1.json 2> gpurun_out/r2b_c1.err; echo "c1 rc=$?"
python bench.py --config C0 --no-cpu > gpurun_out/r2b_c0.json 2> gpurun_out/r2b_c0.err; echo "c0 rc=$?"
python bench.py --config C2 --steps 3 > gpurun_out/r2b_c2.json 2> gpurun_out/r2b_c2.err; echo "c2 rc=$?"
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2b_c4.json 2> gpurun_out/r2b_c4.err; echo "c4 rc=$?"
python bench.py --config C3 --steps 2 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; echo "c3 rc=$?"
for f in c1 c0 c2 c4 c3; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("value"), d["clocks"])
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2b_$f.err").read()[-1500:])
PY
done

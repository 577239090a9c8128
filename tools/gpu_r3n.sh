#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3n_tests.log
tail -3 gpurun_out/r3n_tests.log
line() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d["unit"], round(d["ms_per_step"],4), "ms frac", round(d["roofline"]["frac"],4), "launches", d.get("gpu_launches"))
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1200:])
PY
}
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r3n_c1.json 2> gpurun_out/r3n_c1.err; line r3n_c1
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r3n_c4.json 2> gpurun_out/r3n_c4.err; line r3n_c4

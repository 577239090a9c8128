#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -8 gpurun_out/r2i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2i_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2i_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>&1
grep -c "cub::" gpurun_out/r2i_smoke_launches.csv
python bench.py --config C0 --no-cpu > gpurun_out/r2i_c0.json 2> gpurun_out/r2i_c0.err; echo "c0 rc=$?"
python bench.py --no-cpu > gpurun_out/r2i_c1.json 2> gpurun_out/r2i_c1.err; echo "c1 rc=$?"
for f in c0 c1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2i_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "classes", d["e2e_classes"]["ms_per_step"])
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2i_$f.err").read()[-1500:])
PY
done

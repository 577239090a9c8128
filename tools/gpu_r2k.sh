#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
tail -8 gpurun_out/r2k_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2k_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2k_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>&1
echo "cub kernels in smoke: $(grep -c 'cub::' gpurun_out/r2k_smoke_launches.csv)"
python bench.py > gpurun_out/r2k_c1.json 2> gpurun_out/r2k_c1.err; echo "c1 rc=$?"
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2k_c4.json 2> gpurun_out/r2k_c4.err; echo "c4 rc=$?"
python bench.py --config C0 > gpurun_out/r2k_c0.json 2> gpurun_out/r2k_c0.err; echo "c0 rc=$?"
for f in c1 c4 c0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2k_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "classes", (d.get("e2e_classes") or {}).get("ms_per_step"), "bits", (d.get("e2e_bits") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2k_$f.err").read()[-1500:])
PY
done
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2k_ncu.log 2>&1
echo "ncu rc=$?"

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log
tail -6 gpurun_out/r2s_tests.log
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r2s_c1.json 2> gpurun_out/r2s_c1.err
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2s_c4.json 2> gpurun_out/r2s_c4.err
for f in c1 c4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), d["stages_ms"].get("mc_vertices"))
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2s_$f.err").read()[-1500:])
PY
done
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2s_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2s_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2s_launches.csv 8 | grep "k_mc_vertices\|k_mc_flags\|k_mc_words\|k_mc_emit"

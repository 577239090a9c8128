#!/bin/bash
# round-2 consolidation on one GPU: full GPU suite, smoke, the five bench lines (full default runs), launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3g_tests.log
tail -4 gpurun_out/r3g_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3g_smoke.log
line() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(f, round(d["value"],1), d["unit"], round(d["ms_per_step"],4), "ms frac", round(d["roofline"]["frac"],4), "e2e", e.get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1200:])
PY
}
python bench.py > gpurun_out/r3g_c1.json 2> gpurun_out/r3g_c1.err; line r3g_c1
python bench.py --config C0 > gpurun_out/r3g_c0.json 2> gpurun_out/r3g_c0.err; line r3g_c0
python bench.py --config C4 > gpurun_out/r3g_c4.json 2> gpurun_out/r3g_c4.err; line r3g_c4
python bench.py --config C2 > gpurun_out/r3g_c2.json 2> gpurun_out/r3g_c2.err; line r3g_c2
python bench.py --config C3 > gpurun_out/r3g_c3.json 2> gpurun_out/r3g_c3.err; line r3g_c3
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3g_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r3g_launches_c1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3g_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r3g_launches_c1.csv 8 > gpurun_out/r3g_launches_c1_summary.txt; head -12 gpurun_out/r3g_launches_c1_summary.txt

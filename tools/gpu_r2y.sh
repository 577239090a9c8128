#!/bin/bash
# cheaper byte compare + 3-deep prefetch in k_pack_gap, cheaper cube test in k_mc_flags, grid-stride ambiguous launches:
# full GPU suite, C1 / C4 lines, launch list of the C1 step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2y_tests.log
tail -8 gpurun_out/r2y_tests.log
line() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    k=d["roofline"].get("dominant_kernel",{})
    print(f, round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "| dominant", k.get("kernel","")[:12], round(k.get("us_per_launch",0),1), "us", round(k.get("achieved",0)), "GB/s")
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1200:])
PY
}
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r2y_c1.json 2> gpurun_out/r2y_c1.err; line r2y_c1
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2y_c4.json 2> gpurun_out/r2y_c4.err; line r2y_c4
python bench.py --config C0 --steps 50 --no-e2e --no-cpu > gpurun_out/r2y_c0.json 2> gpurun_out/r2y_c0.err; line r2y_c0
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2y_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2y_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2y_launches.csv 8 | head -24

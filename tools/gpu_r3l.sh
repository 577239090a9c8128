#!/bin/bash
# round-2 final evidence on one GPU: full suite, bench lines (full default runs), launch lists C1 / C4, ncu --set full of one fused C1 step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3l_tests.log
tail -3 gpurun_out/r3l_tests.log
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
python bench.py > gpurun_out/r3l_c1.json 2> gpurun_out/r3l_c1.err; line r3l_c1
python bench.py --config C0 > gpurun_out/r3l_c0.json 2> gpurun_out/r3l_c0.err; line r3l_c0
python bench.py --config C4 > gpurun_out/r3l_c4.json 2> gpurun_out/r3l_c4.err; line r3l_c4
python bench.py --config C2 > gpurun_out/r3l_c2.json 2> gpurun_out/r3l_c2.err; line r3l_c2
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3l_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r3l_launches_c1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3l_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r3l_launches_c1.csv 8 > gpurun_out/r3l_launches_c1_summary.txt
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r3l_launches_c1.csv") if not l.startswith("==")]
rows=list(csv.DictReader(lines))
idx=[i for i,r in enumerate(rows) if r["Kernel Name"].startswith("void k_pack_gap")]
# first kernel of the last complete fused step: two k_pack_flat + k_fill_holes precede k_pack_gap
print("launches", len(rows), "last k_pack_gap at", idx[-3:], "-> skip", idx[-2]-3)
open("gpurun_out/r3l_skip.txt","w").write(str(idx[-2]-3))
PY
SKIP=$(cat gpurun_out/r3l_skip.txt)
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip $SKIP --launch-count 31 -f -o gpurun_out/r3l_step python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3l_ncu_full.log 2>&1; echo "ncu full rc=$?"
python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r3l_plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r3l_launches_c4.csv python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r3l_ncu4.log 2>&1
python tools/ncu_summary.py gpurun_out/r3l_launches_c4.csv 8 > gpurun_out/r3l_launches_c4_summary.txt

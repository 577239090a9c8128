#!/bin/bash
# round-2 GPU call H: layer-segmented z sort (no library sort on the hot path): full suite, C1 / C4 bench A/B, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/r2h_c1.json 2> gpurun_out/r2h_c1.err
T3D_ZSORT_LIBRARY=1 $B > gpurun_out/r2h_c1_cub.json 2> gpurun_out/r2h_c1_cub.err
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2h_c4.json 2> gpurun_out/r2h_c4.err
T3D_ZSORT_LIBRARY=1 python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2h_c4_cub.json 2> gpurun_out/r2h_c4_cub.err
python bench.py --config C3 --steps 3 > gpurun_out/r2h_c3.json 2> gpurun_out/r2h_c3.err
for f in c1 c1_cub c4 c4_cub c3; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), d["stages_ms"].get("canonicalize"))
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/r2h_$f.err").read()[-1500:])
PY
done
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2h_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2h_ncu.log 2>&1
echo "ncu rc=$?"

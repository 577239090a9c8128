#!/bin/bash
# ncu --set full of ONE fused C1 step (32 launches), after the plain command has exited 0
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2v_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r2v_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 1054 --launch-count 34 -f -o gpurun_out/r2v_step python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2v_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2v_step.ncu-rep
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r2v_c1.json 2> gpurun_out/r2v_c1.err
python -c "
import json;d=json.loads(open('gpurun_out/r2v_c1.json').read().strip().splitlines()[-1]);print('c1',round(d['value'],1),round(d['ms_per_step'],4))"

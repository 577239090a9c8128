#!/bin/bash
# ncu --set full of ONE fused C1 step (launches 1064..1092 of `bench.py --steps 2 --warmup 1 --no-e2e --no-cpu`, see the launch list
# of the same command in profiles/r02_launches_c1.csv), after the plain command has exited 0
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3m_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 1064 --launch-count 29 -f -o gpurun_out/r3m_step python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3m_ncu_full.log 2>&1; echo "ncu full rc=$?"

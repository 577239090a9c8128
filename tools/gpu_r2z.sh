#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log
tail -5 gpurun_out/r2z_tests.log
python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r2z_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2z_launches_c4.csv python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r2z_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2z_launches_c4.csv 8 > gpurun_out/r2z_launches_c4_summary.txt; head -40 gpurun_out/r2z_launches_c4_summary.txt

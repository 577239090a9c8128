#!/bin/bash
mkdir -p gpurun_out
python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2o_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2o_launches_c4slab.csv python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2o_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "second_device or fused_single" > gpurun_out/r2o_tests.log 2>&1; tail -2 gpurun_out/r2o_tests.log

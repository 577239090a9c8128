#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3d_tests.log
tail -15 gpurun_out/r3d_tests.log

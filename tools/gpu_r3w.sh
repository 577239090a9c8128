#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voxel.py -m gpu -q > gpurun_out/r3w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3w_tests.log
tail -3 gpurun_out/r3w_tests.log
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r3w_c1.json 2> gpurun_out/r3w_c1.err; python -c "
import json;d=json.loads(open('gpurun_out/r3w_c1.json').read().strip().splitlines()[-1]);k=d['roofline']['dominant_kernel'];print('c1', round(d['ms_per_step'],4), 'dominant', round(k['us_per_launch'],1),'us', round(k['achieved']), 'GB/s frac', round(k['frac'],3))" || tail -5 gpurun_out/r3w_c1.err

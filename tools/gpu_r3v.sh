#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3v_tests.log
tail -3 gpurun_out/r3v_tests.log
for r in 1 2; do python bench.py --steps 20 --no-cpu --no-e2e --no-check > gpurun_out/r3v_c1_$r.json 2> gpurun_out/r3v_c1_$r.err; python -c "
import json;d=json.loads(open('gpurun_out/r3v_c1_$r.json').read().strip().splitlines()[-1]);print('c1', round(d['ms_per_step'],4))" || tail -5 gpurun_out/r3v_c1_$r.err; done

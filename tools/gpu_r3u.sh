#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --no-cpu --no-e2e > gpurun_out/r3u_c1.json 2> gpurun_out/r3u_c1.err; python -c "
import json;d=json.loads(open('gpurun_out/r3u_c1.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],4), d['roofline']['frac'], d['roofline']['frac_of_nominal_8TBs'])" || tail -5 gpurun_out/r3u_c1.err

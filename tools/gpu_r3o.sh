#!/bin/bash
# same box A/B of the two per-layer sort variants
mkdir -p gpurun_out
for rep in 1 2; do for u in 1 8; do
  T3D_ZSORT_U=$u python bench.py --steps 20 --no-cpu --no-e2e --no-check > gpurun_out/r3o_c1_u${u}_$rep.json 2> gpurun_out/r3o_c1_u${u}_$rep.err
  python -c "
import json;d=json.loads(open('gpurun_out/r3o_c1_u${u}_$rep.json').read().strip().splitlines()[-1]);print('C1 U=$u rep $rep', round(d['ms_per_step'],4))"
done; done
for u in 1 8; do
  T3D_ZSORT_U=$u python bench.py --config C4 --steps 5 --no-cpu --no-e2e --no-check > gpurun_out/r3o_c4_u$u.json 2> gpurun_out/r3o_c4_u$u.err
  python -c "
import json;d=json.loads(open('gpurun_out/r3o_c4_u$u.json').read().strip().splitlines()[-1]);print('C4 U=$u', round(d['ms_per_step'],4))"
done

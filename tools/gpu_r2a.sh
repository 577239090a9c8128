#!/bin/bash
# round-2 GPU call A: GPU test suite + variant benches + launch list of the new default path
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
$B > gpurun_out/r2a_new.json 2> gpurun_out/r2a_new.err
T3D_NO_PACK_GAP=1 $B > gpurun_out/r2a_nopg.json 2> gpurun_out/r2a_nopg.err
T3D_NO_FAST_SIGN=1 $B > gpurun_out/r2a_nofs.json 2> gpurun_out/r2a_nofs.err
T3D_NO_PACK_GAP=1 T3D_NO_FAST_SIGN=1 $B > gpurun_out/r2a_old.json 2> gpurun_out/r2a_old.err
T3D_MORPH_ROWS=16 $B > gpurun_out/r2a_my16.json 2> gpurun_out/r2a_my16.err
T3D_PACK_ZC=128 $B > gpurun_out/r2a_zc128.json 2> gpurun_out/r2a_zc128.err
T3D_PACK_ZC=32 $B > gpurun_out/r2a_zc32.json 2> gpurun_out/r2a_zc32.err
python bench.py --shape 512,4096,4096 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2a_c4slab.json 2> gpurun_out/r2a_c4slab.err
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/r2a_tests.log
for f in new nopg nofs old my16 zc128 zc32 c4slab; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", d["stages_ms"])
except Exception as e:
    print("$f FAILED", e)
PY
done

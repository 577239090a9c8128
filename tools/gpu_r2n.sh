#!/bin/bash
mkdir -p gpurun_out
T3D_MORPH_ZTILE=5 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r2n_tests_tiled.log 2>&1; echo "tiled tests rc=$?"; tail -3 gpurun_out/r2n_tests_tiled.log
for t in auto 0 4 8 16 32; do
  if [ "$t" = "auto" ]; then unset T3D_MORPH_ZTILE; else export T3D_MORPH_ZTILE=$t; fi
  python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2n_c4_$t.json 2> gpurun_out/r2n_c4_$t.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_c4_$t.json").read().strip().splitlines()[-1])
    print("c4 ztile=$t", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("c4 ztile=$t FAILED", e); print(open("gpurun_out/r2n_c4_$t.err").read()[-800:])
PY
done
unset T3D_MORPH_ZTILE
python bench.py --steps 20 --no-cpu > gpurun_out/r2n_c1.json 2>gpurun_out/r2n_c1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2n_c1.json').read().strip().splitlines()[-1]); print('c1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'classes', d['e2e_classes']['ms_per_step'], 'bits', d['e2e_bits']['ms_per_step'])"

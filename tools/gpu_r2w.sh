#!/bin/bash
# one-pass opening+closing (k_morph4): full GPU suite, then A/B against the four-launch chain and a small tile sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2w_tests.log
tail -8 gpurun_out/r2w_tests.log
line() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "smooth", round(d["stages_ms"].get("smooth"),4))
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/%s.err"%f).read()[-1200:])
PY
}
python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r2w_c1.json 2> gpurun_out/r2w_c1.err; line r2w_c1
T3D_NO_MORPH4=1 python bench.py --steps 20 --no-cpu --no-e2e > gpurun_out/r2w_c1_chain.json 2> gpurun_out/r2w_c1_chain.err; line r2w_c1_chain
python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2w_c4.json 2> gpurun_out/r2w_c4.err; line r2w_c4
T3D_NO_MORPH4=1 python bench.py --config C4 --steps 5 --no-e2e --no-cpu > gpurun_out/r2w_c4_chain.json 2> gpurun_out/r2w_c4_chain.err; line r2w_c4_chain
for cfg in "56 32" "56 64" "120 32" "120 64" "24 64"; do set -- $cfg
  T3D_MORPH4_TY=$1 T3D_MORPH4_ZC=$2 python bench.py --steps 20 --no-cpu --no-e2e --no-check > gpurun_out/r2w_c1_ty$1_zc$2.json 2> gpurun_out/r2w_c1_ty$1_zc$2.err; line r2w_c1_ty$1_zc$2
done
for cfg in "24 64" "24 32"; do set -- $cfg
  T3D_MORPH4_TY=$1 T3D_MORPH4_ZC=$2 python bench.py --config C4 --steps 5 --no-cpu --no-e2e --no-check > gpurun_out/r2w_c4_ty$1_zc$2.json 2> gpurun_out/r2w_c4_ty$1_zc$2.err; line r2w_c4_ty$1_zc$2
done

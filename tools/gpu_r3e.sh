#!/bin/bash
# N=2: sharded step with the pre-filled slab mode, C1 and C4-shaped slabs, against the plain mode (T3D_NO_SLAB_PACK_GAP)
mkdir -p gpurun_out
run() { # name, extra env, config, steps
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --config $3 --steps $4 --warmup 3 --no-cpu --no-e2e > gpurun_out/r3e_$1.json 2> gpurun_out/r3e_$1.err
  python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r3e_%s.json"%f).read().strip().splitlines()[-1])
    c=d.get("sharded_check") or {}
    print(f, round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms frac", round(d["roofline"]["frac"],4), "check ok" if c.get("ok") else c)
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/r3e_%s.err"%f).read()[-1500:])
PY
}
run c1 "X=1" C1 20
run c1_plain "T3D_NO_SLAB_PACK_GAP=1" C1 20
run c4 "X=1" C4 5
run c4_plain "T3D_NO_SLAB_PACK_GAP=1" C4 5

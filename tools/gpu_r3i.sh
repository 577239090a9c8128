#!/bin/bash
# N=8: C1 (weak scaling, 8 x 512 x 1024 x 1024) and C4 (4096^3) lines, sharded step with the pre-filled slab mode
mkdir -p gpurun_out
run() { # name, config, steps, extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --config $2 --steps $3 --warmup 3 --no-cpu $4 > gpurun_out/r3j_$1.json 2> gpurun_out/r3j_$1.err
  python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r3j_%s.json"%f).read().strip().splitlines()[-1])
    c=d.get("sharded_check") or {}
    e=d.get("e2e") or {}
    print(f, round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms frac", round(d["roofline"]["frac"],4), "e2e", e.get("value"), "check ok" if c.get("ok") else c, d["config"].get("partition"))
except Exception as e:
    print(f, "FAILED", e); print(open("gpurun_out/r3j_%s.err"%f).read()[-1500:])
PY
}
run c1 C1 20 ""
run c4 C4 5 "--no-e2e"

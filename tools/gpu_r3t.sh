#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 10 --no-cpu --no-e2e > gpurun_out/r3t_c1_n2.json 2> gpurun_out/r3t_c1_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r3t_c1_n2.json").read().strip().splitlines()[-1])
    c=d.get("sharded_check") or {}
    print("c1 n2", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "check ok" if c.get("ok") else c)
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r3t_c1_n2.err").read()[-1500:])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --no-cpu > gpurun_out/r3q_c1_n4.json 2> gpurun_out/r3q_c1_n4.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r3q_c1_n4.json").read().strip().splitlines()[-1])
    c=d.get("sharded_check") or {}
    print("c1 n4", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("value"), "check ok" if c.get("ok") else c, d["config"].get("partition"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r3q_c1_n4.err").read()[-1500:])
PY

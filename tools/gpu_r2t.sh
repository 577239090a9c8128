#!/bin/bash
# N=2: C4-shaped slabs (4096 x 4096 slices), balanced partition (area-scaled cost model) against the equal split
mkdir -p gpurun_out
for mode in bal eq; do
  extra=""; [ $mode = eq ] && extra="--no-balance"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config C4 --steps 5 --warmup 3 --no-cpu --no-e2e $extra > gpurun_out/r2t_c4_n2_$mode.json 2> gpurun_out/r2t_c4_n2_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2t_c4_n2_$mode.json").read().strip().splitlines()[-1])
    print("$mode", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", d["config"].get("partition"), d.get("sharded_check"))
except Exception as e:
    print("$mode FAILED", e); print(open("gpurun_out/r2t_c4_n2_$mode.err").read()[-1500:])
PY
done

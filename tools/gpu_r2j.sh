#!/bin/bash
# round-2 multi-GPU call: N ranks (argument), torchrun as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR tools/sharded_check.py > gpurun_out/r2j_check_n$N.log 2>&1; echo "sharded_check rc=$?"; tail -12 gpurun_out/r2j_check_n$N.log
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2j_c1_n$N.json 2> gpurun_out/r2j_c1_n$N.err; echo "c1 rc=$?"
$TR bench.py --gpus $N --steps 20 --warmup 3 --no-balance --no-e2e --no-check > gpurun_out/r2j_c1_n${N}_equal.json 2> gpurun_out/r2j_c1_n${N}_equal.err; echo "c1 equal rc=$?"
$TR bench.py --gpus $N --config C3 --steps 3 > gpurun_out/r2j_c3_n$N.json 2> gpurun_out/r2j_c3_n$N.err; echo "c3 rc=$?"
$TR bench.py --gpus $N --config C2 --steps 3 > gpurun_out/r2j_c2_n$N.json 2> gpurun_out/r2j_c2_n$N.err; echo "c2 rc=$?"
if [ "$N" = "8" ]; then
  $TR bench.py --gpus $N --config C4 --steps 5 --no-e2e > gpurun_out/r2j_c4_n$N.json 2> gpurun_out/r2j_c4_n$N.err; echo "c4 rc=$?"
fi
for f in c1_n$N c1_n${N}_equal c3_n$N c2_n$N c4_n$N; do python - <<PY
import json, os
p="gpurun_out/r2j_$f.json"
if os.path.exists(p):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print("$f", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms", "frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("value"), d["config"].get("partition"), (d.get("sharded_check") or {}).get("ok"), d["clocks"]["samples"])
    except Exception as e:
        print("$f FAILED", e); print(open("gpurun_out/r2j_$f.err").read()[-2500:])
PY
done

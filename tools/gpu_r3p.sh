#!/bin/bash
# final N=2 evidence: full GPU suite on a 2-GPU box (two-device test included), C1 line with e2e, sharded check
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3p_tests.log
tail -3 gpurun_out/r3p_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --no-cpu > gpurun_out/r3p_c1_n2.json 2> gpurun_out/r3p_c1_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r3p_c1_n2.json").read().strip().splitlines()[-1])
    c=d.get("sharded_check") or {}
    print("c1 n2", round(d["value"],1), "Gvox/s", round(d["ms_per_step"],4), "ms frac", round(d["roofline"]["frac"],4), "e2e", (d.get("e2e") or {}).get("value"), "check ok" if c.get("ok") else c, d["config"].get("partition"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r3p_c1_n2.err").read()[-1500:])
PY
grep -c "NCCL INFO" gpurun_out/r3p_c1_n2.err

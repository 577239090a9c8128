#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/sdf_time.py <<'PY'
import torch, time, sys
sys.path.insert(0, ".")
import bench
from tomography_3d_reconstructor_b200 import engine, edt
dev = torch.device("cuda", 0)
Z, H, W = 512, 1024, 1024
masks = bench.make_phantom_u8(Z, H, W, 0, Z, dev)
dv = engine.smooth(engine.pack_and_close(masks, 200, True), 3, True)
del masks
samp = (6.0 / Z, 95.03 / H, 143.1 / W)
for _ in range(2):
    out = edt.signed_distance(dv, samp); torch.cuda.synchronize()
PY
ncu --set full --import-source on --clock-control none -k regex:k_sdf_envelope -s 2 -c 2 -o gpurun_out/r2g_sdf python /tmp/sdf_time.py > gpurun_out/r2g_ncu.log 2>&1
ls -la gpurun_out/r2g_sdf.ncu-rep

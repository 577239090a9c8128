#!/bin/bash
mkdir -p gpurun_out
python tools/slab_mode_probe.py 256 4096 4096 > gpurun_out/r3f_pre.log 2>&1; tail -3 gpurun_out/r3f_pre.log
T3D_NO_SLAB_PACK_GAP=1 python tools/slab_mode_probe.py 256 4096 4096 > gpurun_out/r3f_plain.log 2>&1; tail -3 gpurun_out/r3f_plain.log
T3D_STAGE_EVENTS=1 python tools/slab_mode_probe.py 256 4096 4096 2>&1 | grep "t3d stages" | tail -1
T3D_STAGE_EVENTS=1 T3D_NO_SLAB_PACK_GAP=1 python tools/slab_mode_probe.py 256 4096 4096 2>&1 | grep "t3d stages" | tail -1

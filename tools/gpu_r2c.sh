#!/bin/bash
# round-2 GPU call C: full GPU suite, unmodified orchestrator on the drop-in, compute-sanitizer (memcheck + racecheck), C1 launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -15 gpurun_out/r2c_tests.log
if [ -d oracle/_ref/reference_checkout ]; then
  python tools/run_orchestrator.py --reference oracle/_ref/reference_checkout --log gpurun_out/r2c_orchestrator.log > gpurun_out/r2c_orch.out 2>&1; echo "orchestrator rc=$?"
  tail -12 gpurun_out/r2c_orch.out
fi
SAN="tests/test_gpu_voxel.py tests/test_gpu_surface.py tests/test_gpu_pipeline.py tests/test_gpu_edt.py tests/test_gpu_export.py tests/test_gpu_dropin.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2c_memcheck.log python -m pytest $SAN -m gpu -q -x > gpurun_out/r2c_memcheck.out 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r2c_memcheck.out; tail -5 gpurun_out/r2c_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2c_racecheck.log python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_voxel.py -m gpu -q -x -k "not sharded_sdf" > gpurun_out/r2c_racecheck.out 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2c_racecheck.out; tail -5 gpurun_out/r2c_racecheck.log
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"

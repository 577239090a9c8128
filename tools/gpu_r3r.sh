#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3r_tests.log
tail -4 gpurun_out/r3r_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

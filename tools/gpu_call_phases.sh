#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "dit_" > gpurun_out/s7_pytest.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/s7_pytest.log
timeout 300 python tools/prof_dit_block.py 1024 > gpurun_out/s7_prof.log 2>&1; cat gpurun_out/s7_prof.log

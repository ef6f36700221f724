#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "strided_conv or super_resolution or cascade or query_blocks or c7" > gpurun_out/s9_pytest.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/s9_pytest.log

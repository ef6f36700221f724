#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/s8_pytest.log

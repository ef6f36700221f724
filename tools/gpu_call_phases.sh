#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "c7 or lora or checkpoint or dropin" > gpurun_out/s3_pytest_new.log 2>&1
echo "new tests rc=$?"; tail -15 gpurun_out/s3_pytest_new.log
cp build/libxdb200_inst.so xdiffusion_b200/libxdb200.so
for a in "1024 1" "128 1" "128 4" "512 2"; do timeout 120 python tools/prof_dit_phases.py $a; done > gpurun_out/s3_phases.log 2>&1
tail -5 gpurun_out/s3_phases.log

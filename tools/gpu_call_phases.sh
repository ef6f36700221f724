#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "c7 or lora or checkpoint or dit_ or rows_are_independent or full_length" > gpurun_out/s4_pytest_new.log 2>&1
echo "new tests rc=$?"; tail -15 gpurun_out/s4_pytest_new.log
timeout 300 python tools/prof_dit_block.py 1024 > gpurun_out/s4_prof.log 2>&1; cat gpurun_out/s4_prof.log
cp build/libxdb200_inst.so xdiffusion_b200/libxdb200.so
for a in "1024 1"; do timeout 120 python tools/prof_dit_phases.py $a; done > gpurun_out/s4_phases.log 2>&1
tail -5 gpurun_out/s4_phases.log

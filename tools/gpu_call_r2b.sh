#!/bin/bash
mkdir -p gpurun_out
python tools/prof_edm.py > gpurun_out/z_edm.log 2>&1; cat gpurun_out/z_edm.log | tail -4
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/z_r2b python tools/ncu_targets.py "r2b" > gpurun_out/z_ncu_r2b.log 2>&1
tail -2 gpurun_out/z_ncu_r2b.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/z_launches_unet.csv python bench.py --workload unet --steps 1 --warmup 1 --sampling-steps 25 --no-cpu > gpurun_out/z_ncu_unet.log 2>&1
ls -la gpurun_out/z_*

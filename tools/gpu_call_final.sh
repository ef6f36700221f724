#!/bin/bash
# bench (no profiler) -> launch list -> ncu --set full of the fused DiT kernels
mkdir -p gpurun_out
(time python bench.py --steps 3 --warmup 3) > gpurun_out/f1_bench.log 2>&1
tail -4 gpurun_out/f1_bench.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/f1_launches_dit.csv python bench.py --workload dit --steps 1 --warmup 1 --sampling-steps 20 --no-cpu > gpurun_out/f1_ncu_dit.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/f1_dit_fused python tools/ncu_targets.py "dit " > gpurun_out/f1_ncu_full.log 2>&1
tail -2 gpurun_out/f1_ncu_full.log
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/f1_attn64 python tools/ncu_targets.py "T=64" > gpurun_out/f1_ncu_attn64.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; "$@" > gpurun_out/u_$tag.log 2>&1; grep '^{' gpurun_out/u_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],2), d['unit'], round(d['ms_per_timestep'],4), 'ms/timestep', d['launches_per_timestep'], 'launches')" || tail -5 gpurun_out/u_$tag.log; }
run unet1 timeout 900 python bench.py --workload unet --steps 1 --warmup 1 --no-cpu
XDB200_UNET_STREAMS=2 run unet2 timeout 900 python bench.py --workload unet --steps 1 --warmup 1 --no-cpu
XDB200_UNET_STREAMS=2 timeout 600 python -m pytest tests/test_e2e_gpu.py tests/test_fulllen_gpu.py -x -q -k "c1 or c3 or c6 or unet" > gpurun_out/u_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/u_tests.log

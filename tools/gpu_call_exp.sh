#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; "$@" > gpurun_out/x_$tag.log 2>&1; grep '^{' gpurun_out/x_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],1), 'img/s', round(d['ms_per_timestep'],4), 'ms/timestep', d['launches_per_timestep'], 'launches', d['roofline']['per_shape'])"; }
run dit1024 timeout 600 python bench.py --workload dit --steps 2 --warmup 2 --no-cpu
run dit128_pdl2 timeout 600 python bench.py --workload dit --batch 128 --steps 2 --warmup 2 --no-cpu
XDB200_PDL=1 run dit128_pdl1 timeout 600 python bench.py --workload dit --batch 128 --steps 2 --warmup 2 --no-cpu
run dit256 timeout 600 python bench.py --workload dit --batch 256 --steps 2 --warmup 2 --no-cpu
run dit512 timeout 600 python bench.py --workload dit --batch 512 --steps 2 --warmup 2 --no-cpu
XDB200_DIT_FUSED_MIN_ROWS=0 run dit512_fused timeout 600 python bench.py --workload dit --batch 512 --steps 2 --warmup 2 --no-cpu

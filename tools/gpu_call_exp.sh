#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -x -q -k "dit_ or c2 or bit_identical or rows_are_independent" > gpurun_out/x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/x_tests.log
run() { tag=$1; shift; "$@" > gpurun_out/x_$tag.log 2>&1; grep '^{' gpurun_out/x_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],1), 'img/s', round(d['ms_per_timestep'],4), 'ms/timestep', d['launches_per_timestep'], 'launches')"; }
run prefetch timeout 600 python bench.py --workload dit --steps 2 --warmup 2 --no-cpu
XDB200_DIT_PREFETCH=0 run noprefetch timeout 600 python bench.py --workload dit --steps 2 --warmup 2 --no-cpu

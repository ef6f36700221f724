#!/bin/bash
# one gpurun call: validate + time the fused DiT kernels (hidden-split MLP), UNet launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "dit_" -x -q > gpurun_out/s2_pytest_mlp.log 2>&1
echo "dit kernel tests rc=$?"; tail -3 gpurun_out/s2_pytest_mlp.log
for b in 128 256 512 1024; do timeout 300 python tools/prof_dit_block.py $b; done > gpurun_out/s2_prof.log 2>&1
cat gpurun_out/s2_prof.log
for b in 128 256 512 1024; do
  timeout 600 python bench.py --workload dit --batch $b --steps 2 --warmup 1 --no-cpu > gpurun_out/s2_bench_b$b.log 2>&1
  grep '^{' gpurun_out/s2_bench_b$b.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dit', $b, d['value'], d['ms_per_timestep'], d['roofline']['per_shape'])"
done
XDB200_DIT_MLP_SPLIT=1 timeout 600 python bench.py --workload dit --batch 128 --steps 2 --warmup 1 --no-cpu > gpurun_out/s2_bench_b128_nosplit.log 2>&1
grep '^{' gpurun_out/s2_bench_b128_nosplit.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dit nosplit 128', d['value'], d['ms_per_timestep'])"
XDB200_DIT_FUSED=0 timeout 600 python bench.py --workload dit --batch 128 --steps 2 --warmup 1 --no-cpu > gpurun_out/s2_bench_b128_unfused.log 2>&1
grep '^{' gpurun_out/s2_bench_b128_unfused.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dit unfused 128', d['value'], d['ms_per_timestep'])"
for pdl in 2 1; do
  XDB200_PDL=$pdl timeout 600 python bench.py --workload unet --steps 1 --warmup 1 --no-cpu > gpurun_out/s2_unet_pdl$pdl.log 2>&1
  grep '^{' gpurun_out/s2_unet_pdl$pdl.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unet pdl=$pdl', d['value'], d['ms_per_timestep'])"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest_all.log 2>&1
echo "all tests rc=$?"; tail -4 gpurun_out/s2_pytest_all.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/s2_launches_unet.csv python bench.py --workload unet --steps 1 --warmup 1 --sampling-steps 12 --no-cpu > gpurun_out/s2_ncu_unet.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s2_launches_dit128.csv python bench.py --workload dit --batch 128 --steps 1 --warmup 1 --sampling-steps 20 --no-cpu > gpurun_out/s2_ncu_dit128.log 2>&1
echo done

#!/bin/bash
mkdir -p gpurun_out
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3) > gpurun_out/n2_bench.log 2>&1
tail -3 gpurun_out/n2_bench.log | cut -c1-3000
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1) > gpurun_out/n2_ref.log 2>&1
tail -3 gpurun_out/n2_ref.log | cut -c1-1200

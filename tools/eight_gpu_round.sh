#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/bench8.err
grep '^{' gpurun_out/r02_bench_8gpu.json | tail -1 | cut -c1-300; tail -2 gpurun_out/bench8.err

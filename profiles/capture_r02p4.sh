#!/bin/bash
# r02p4: four ranks on one box
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02p4_bench_4gpu.json 2> gpurun_out/r02p4_bench_4gpu.err
tail -c 400 gpurun_out/r02p4_bench_4gpu.err
head -c 300 gpurun_out/r02p4_bench_4gpu.json

#!/bin/bash
# r02p8: eight ranks on one box (both arms)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02p8_bench_8gpu.json 2> gpurun_out/r02p8_bench_8gpu.err
tail -c 400 gpurun_out/r02p8_bench_8gpu.err
head -c 300 gpurun_out/r02p8_bench_8gpu.json

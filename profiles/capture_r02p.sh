#!/bin/bash
# r02p: two ranks on one box (torchrun, NCCL barrier, NUMA binding of the feeder threads), reference arm next to it
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02p_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|^CPU\(s\)|Model name" >> gpurun_out/r02p_topo.txt
cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02p_topo.txt 2>&1
tail -8 gpurun_out/r02p_topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02p_bench_2gpu.json 2> gpurun_out/r02p_bench_2gpu.err
tail -c 600 gpurun_out/r02p_bench_2gpu.err
head -c 300 gpurun_out/r02p_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02p_bench_ref_2gpu.json 2> gpurun_out/r02p_bench_ref_2gpu.err
tail -c 300 gpurun_out/r02p_bench_ref_2gpu.err
head -c 300 gpurun_out/r02p_bench_ref_2gpu.json
PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02p_probe.txt 2>&1
cat gpurun_out/r02p_probe.txt

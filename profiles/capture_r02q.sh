#!/bin/bash
# r02q: cost-only last round, marginalisation-mode early exit; tests, probe, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_gpu_tests.log 2>&1
tail -4 gpurun_out/r02q_gpu_tests.log
grep -q " passed" gpurun_out/r02q_gpu_tests.log || tail -60 gpurun_out/r02q_gpu_tests.log
PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02q_probe.txt 2>&1
cat gpurun_out/r02q_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
tail -c 800 gpurun_out/r02q_bench.err
head -c 400 gpurun_out/r02q_bench.json

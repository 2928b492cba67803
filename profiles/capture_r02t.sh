#!/bin/bash
# r02t: partial sums in the dot-product chains (marg_tri, lin_small, solve symv, prior_setup): tests, probe, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_gpu_tests.log 2>&1
tail -4 gpurun_out/r02t_gpu_tests.log
grep -q " passed" gpurun_out/r02t_gpu_tests.log || tail -60 gpurun_out/r02t_gpu_tests.log
PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02t_probe.txt 2>&1
PROBE_COPIES=48 PROBE_CONFIG=4 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02t_probe.txt 2>&1
cat gpurun_out/r02t_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
tail -c 800 gpurun_out/r02t_bench.err
head -c 400 gpurun_out/r02t_bench.json

#!/bin/bash
# r02u: marg_tri reorganised (T lanes per row, 9 barriers per column instead of 15): marginalisation parity tests, probes at 128 / 192 / 256 threads
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_gpu_tests.log 2>&1
tail -4 gpurun_out/r02u_gpu_tests.log
grep -q " passed" gpurun_out/r02u_gpu_tests.log || tail -60 gpurun_out/r02u_gpu_tests.log
PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so profiles/var_tri192.so profiles/var_tri256.so > gpurun_out/r02u_probe.txt 2>&1
PROBE_COPIES=48 PROBE_CONFIG=4 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so profiles/var_tri256.so >> gpurun_out/r02u_probe.txt 2>&1
cat gpurun_out/r02u_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 --overlap-lk > gpurun_out/r02u_bench_overlap.json 2> gpurun_out/r02u_bench_overlap.err
tail -c 300 gpurun_out/r02u_bench_overlap.err
head -c 300 gpurun_out/r02u_bench_overlap.json

#!/bin/bash
# r02w: block-wide Jacobi again in marg_prep, marg_tri update loops in batches of four: tests + probe
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_gpu_tests.log 2>&1
tail -4 gpurun_out/r02w_gpu_tests.log
grep -q " passed" gpurun_out/r02w_gpu_tests.log || tail -60 gpurun_out/r02w_gpu_tests.log
PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02w_probe.txt 2>&1
PROBE_COPIES=48 PROBE_CONFIG=4 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02w_probe.txt 2>&1
cat gpurun_out/r02w_probe.txt

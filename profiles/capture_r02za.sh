#!/bin/bash
# r02za: landmark-block descriptor + code-word reads in lin_vis_lm: tests + probe
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02za_gpu_tests.log 2>&1
tail -4 gpurun_out/r02za_gpu_tests.log
grep -q " passed" gpurun_out/r02za_gpu_tests.log || tail -60 gpurun_out/r02za_gpu_tests.log
PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02za_probe.txt 2>&1
PROBE_COPIES=48 PROBE_CONFIG=4 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02za_probe.txt 2>&1
cat gpurun_out/r02za_probe.txt

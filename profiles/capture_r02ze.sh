#!/bin/bash
# r02ze: the round's last state (pair chunks of 64 records, asm_pairs + pair_reduce): GPU tests + probe
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02ze_gpu_tests.log 2>&1
tail -2 gpurun_out/r02ze_gpu_tests.log
PROBE_COPIES=48 timeout 200 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02ze_probe.txt 2>&1
cat gpurun_out/r02ze_probe.txt

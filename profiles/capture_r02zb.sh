#!/bin/bash
# r02zb: LK border staging word-wise: tests + LK probe
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zb_gpu_tests.log 2>&1
tail -4 gpurun_out/r02zb_gpu_tests.log
grep -q " passed" gpurun_out/r02zb_gpu_tests.log || tail -60 gpurun_out/r02zb_gpu_tests.log
timeout 200 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02zb_lk_probe.txt 2>&1
cat gpurun_out/r02zb_lk_probe.txt

#!/bin/bash
# r02y: where the upload time of the end-to-end step goes
set -x
mkdir -p gpurun_out
VIWB_TIMING=1 timeout 600 python profiles/h2d_probe.py > gpurun_out/r02y_h2d_probe.txt 2> gpurun_out/r02y_h2d_probe.err
cat gpurun_out/r02y_h2d_probe.txt
grep "build B=1776" gpurun_out/r02y_h2d_probe.err | tail -3
tail -3 gpurun_out/r02y_h2d_probe.err

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1
tail -5 gpurun_out/r02b_gpu_tests.log
bash profiles/capture_r02b.sh

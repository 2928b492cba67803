#!/bin/bash
# r02s: upload not waited for in batch_build (results staged in their own region of the pinned slab): tests, e2e lane sweep
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_gpu_tests.log 2>&1
tail -4 gpurun_out/r02s_gpu_tests.log
grep -q " passed" gpurun_out/r02s_gpu_tests.log || tail -60 gpurun_out/r02s_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --e2e-lanes-sweep 3,6,8 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
tail -c 800 gpurun_out/r02s_bench.err
head -c 400 gpurun_out/r02s_bench.json
VIWB_TIMING=1 timeout 300 python profiles/e2e_probe.py 2> gpurun_out/r02s_e2e_probe.txt
grep -B4 "B=1024 call" gpurun_out/r02s_e2e_probe.txt | tail -5

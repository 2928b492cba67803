#!/bin/bash
# r02o: packed wire format of the visual table, lin_small 4 blocks/SM, marg_tri 128 threads: tests, probes (incl. one window), bench + e2e lane sweep
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_gpu_tests.log 2>&1
tail -4 gpurun_out/r02o_gpu_tests.log
grep -q " passed" gpurun_out/r02o_gpu_tests.log || tail -60 gpurun_out/r02o_gpu_tests.log
PROBE_COPIES=48 timeout 900 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so profiles/var_lsm5.so profiles/var_lsm6.so profiles/var_tri64.so profiles/var_tri96.so > gpurun_out/r02o_probe.txt 2>&1
PROBE_DISTINCT=1 PROBE_COPIES=1 timeout 300 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02o_probe.txt 2>&1
cat gpurun_out/r02o_probe.txt
VIWB_TIMING=1 timeout 300 python profiles/e2e_probe.py 2> gpurun_out/r02o_e2e_probe.txt
grep "B=1024" gpurun_out/r02o_e2e_probe.txt | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --e2e-lanes-sweep 6,8 > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
tail -c 800 gpurun_out/r02o_bench.err
head -c 400 gpurun_out/r02o_bench.json

#!/bin/bash
# r02zc: pair_win (asm_pairs + pair_reduce fused per window, chunk products in shared memory): tests, probe against the two-kernel path, bench line, ncu of the new kernel
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zc_gpu_tests.log 2>&1
tail -3 gpurun_out/r02zc_gpu_tests.log
grep -q " passed" gpurun_out/r02zc_gpu_tests.log || tail -60 gpurun_out/r02zc_gpu_tests.log
PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02zc_probe.txt 2>&1
VIWB_NO_PAIR_WIN=1 PROBE_COPIES=48 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02zc_probe.txt 2>&1
PROBE_COPIES=48 PROBE_CONFIG=4 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02zc_probe.txt 2>&1
cat gpurun_out/r02zc_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02zc_bench.json 2> gpurun_out/r02zc_bench.err
tail -c 300 gpurun_out/r02zc_bench.err
head -c 300 gpurun_out/r02zc_bench.json
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'^pair_win_kernel' -o gpurun_out/r02zc_pair_win python profiles/ncu_target.py --iters 1 --no-lk > gpurun_out/r02zc_ncu.log 2>&1
tail -2 gpurun_out/r02zc_ncu.log
xz -T0 -3 gpurun_out/r02zc_pair_win.ncu-rep

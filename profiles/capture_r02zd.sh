#!/bin/bash
# r02zd: pair chunks of 64 records, pair_win active on the bench window (109 KB of chunk products per block): tests, probe with and without, bench line, ncu of pair_win
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02zd_gpu_tests.log 2>&1
tail -3 gpurun_out/r02zd_gpu_tests.log
grep -q " passed" gpurun_out/r02zd_gpu_tests.log || tail -40 gpurun_out/r02zd_gpu_tests.log
PROBE_COPIES=48 timeout 300 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02zd_probe.txt 2>&1
VIWB_NO_PAIR_WIN=1 PROBE_COPIES=48 timeout 300 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02zd_probe.txt 2>&1
cat gpurun_out/r02zd_probe.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02zd_bench.json 2> gpurun_out/r02zd_bench.err
tail -c 300 gpurun_out/r02zd_bench.err
head -c 300 gpurun_out/r02zd_bench.json
timeout 200 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'^pair_win_kernel' -c 2 -o gpurun_out/r02zd_pair_win python profiles/ncu_target.py --iters 1 --no-lk > gpurun_out/r02zd_ncu.log 2>&1
tail -2 gpurun_out/r02zd_ncu.log
xz -T0 -3 gpurun_out/r02zd_pair_win.ncu-rep || true

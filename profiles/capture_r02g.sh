#!/bin/bash
# r02g: TMA isolation test, pipelined TMA staging in lk_track (patch of the next level / search region of this level requested ahead), occupancy variants
set -x
mkdir -p gpurun_out
for v in "1 0" "2 0" "2 1" "3 0" "3 1"; do timeout 60 ./profiles/micro/tma_test $v; done > gpurun_out/r02g_tma_test.txt 2>&1
cat gpurun_out/r02g_tma_test.txt
timeout 200 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so profiles/variants_lk6.so profiles/variants_lk8.so > gpurun_out/r02g_lk_probe.txt 2>&1
if ! grep -q "tick" gpurun_out/r02g_lk_probe.txt; then tail -5 gpurun_out/r02g_lk_probe.txt; echo "TMA PATH FAILED -- continuing without it"; export VIWB_LK_NO_TMA=1; fi
VIWB_LK_NO_TMA=1 timeout 300 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so profiles/variants_lk6.so profiles/variants_lk8.so >> gpurun_out/r02g_lk_probe.txt 2>&1
cat gpurun_out/r02g_lk_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gpu_tests.log 2>&1
tail -4 gpurun_out/r02g_gpu_tests.log
grep -q " passed" gpurun_out/r02g_gpu_tests.log || tail -60 gpurun_out/r02g_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
tail -c 600 gpurun_out/r02g_bench.err
head -c 1200 gpurun_out/r02g_bench.json
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(lk_track_tasks)_kernel' \
    -o gpurun_out/r02g_lk python profiles/ncu_target.py --iters 1 > gpurun_out/r02g_ncu.log 2>&1
tail -3 gpurun_out/r02g_ncu.log
xz -T0 -3 gpurun_out/r02g_lk.ncu-rep
ls -la gpurun_out/

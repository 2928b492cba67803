#!/bin/bash
# r02l: LK with 16-byte aligned 48 x 32 TMA tiles; wide fused path (marginalisation + C3 / C4 solver); tests; probes; bench; ncu
set -x
mkdir -p gpurun_out
for v in "0 48 32 0 0 752 16 70" "0 48 32 0 0 752 8 70" "0 48 32 0 0 94 48 70"; do timeout 60 ./profiles/micro/tma_probe $v; done > gpurun_out/r02l_tma_probe.txt 2>&1
cat gpurun_out/r02l_tma_probe.txt
timeout 200 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so profiles/variants_lk5.so > gpurun_out/r02l_lk_probe.txt 2>&1
if ! grep -q "tick" gpurun_out/r02l_lk_probe.txt; then tail -5 gpurun_out/r02l_lk_probe.txt; echo "TMA PATH FAILED -- continuing without it"; export VIWB_LK_NO_TMA=1; fi
VIWB_LK_NO_TMA=1 timeout 300 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02l_lk_probe.txt 2>&1
cat gpurun_out/r02l_lk_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gpu_tests.log 2>&1
tail -4 gpurun_out/r02l_gpu_tests.log
grep -q " passed" gpurun_out/r02l_gpu_tests.log || tail -60 gpurun_out/r02l_gpu_tests.log
PROBE_COPIES=16 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02l_probe.txt 2>&1
PROBE_COPIES=16 PROBE_CONFIG=4 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02l_probe.txt 2>&1
PROBE_COPIES=16 PROBE_CONFIG=4 VIWB_NO_FUSED=1 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02l_probe.txt 2>&1
cat gpurun_out/r02l_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
tail -c 600 gpurun_out/r02l_bench.err
head -c 1200 gpurun_out/r02l_bench.json
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(lk_track_tasks|lin_vis_lm_wide|asm_pairs_wide|pair_reduce|marg_prep)_kernel' \
    -o gpurun_out/r02l_new python profiles/ncu_target.py --iters 1 > gpurun_out/r02l_ncu.log 2>&1
tail -3 gpurun_out/r02l_ncu.log
xz -T0 -3 gpurun_out/r02l_new.ncu-rep
ls -la gpurun_out/

#!/bin/bash
# r02f: TMA descriptors in global memory; solve without the third W pass; everything else as r02e
set -x
mkdir -p gpurun_out
timeout 300 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02f_lk_probe.txt 2>&1
if ! grep -q "tick" gpurun_out/r02f_lk_probe.txt; then tail -5 gpurun_out/r02f_lk_probe.txt; echo "TMA PATH FAILED -- continuing without it"; export VIWB_LK_NO_TMA=1; fi
VIWB_LK_NO_TMA=1 timeout 300 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02f_lk_probe.txt 2>&1
cat gpurun_out/r02f_lk_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.log 2>&1
tail -4 gpurun_out/r02f_gpu_tests.log
grep -q " passed" gpurun_out/r02f_gpu_tests.log || tail -60 gpurun_out/r02f_gpu_tests.log
PROBE_COPIES=16 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02f_probe.txt 2>&1
cat gpurun_out/r02f_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -c 600 gpurun_out/r02f_bench.err
head -c 1500 gpurun_out/r02f_bench.json
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(marg_prep|marg_eig|lin_vis_lm|pair_reduce|asm_pairs|lk_track_tasks|solve)_kernel' \
    -o gpurun_out/r02f_new python profiles/ncu_target.py --iters 1 > gpurun_out/r02f_ncu.log 2>&1
tail -3 gpurun_out/r02f_ncu.log
xz -T0 -3 gpurun_out/r02f_new.ncu-rep
ls -la gpurun_out/

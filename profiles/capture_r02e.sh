#!/bin/bash
# r02e: marg split (prep + tql2 eig), lin_vis_lm with smem metadata (3 / 4 blocks per SM), pair_reduce, TMA + dp2a LK
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gpu_tests.log 2>&1
tail -4 gpurun_out/r02e_gpu_tests.log
grep -q " passed" gpurun_out/r02e_gpu_tests.log || tail -80 gpurun_out/r02e_gpu_tests.log
PROBE_COPIES=16 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02e_probe.txt 2>&1
cat gpurun_out/r02e_probe.txt
timeout 600 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02e_lk_probe.txt 2>&1
VIWB_LK_NO_TMA=1 timeout 600 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02e_lk_probe.txt 2>&1
cat gpurun_out/r02e_lk_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -c 600 gpurun_out/r02e_bench.err
head -c 1500 gpurun_out/r02e_bench.json
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(marg_prep|marg_eig|lin_vis_lm|pair_reduce|asm_pairs|lk_track_tasks|solve)_kernel' \
    -o gpurun_out/r02e_new python profiles/ncu_target.py --iters 1 > gpurun_out/r02e_ncu.log 2>&1
tail -3 gpurun_out/r02e_ncu.log
xz -T0 -3 gpurun_out/r02e_new.ncu-rep
ls -la gpurun_out/

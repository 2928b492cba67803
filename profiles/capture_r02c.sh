#!/bin/bash
# r02c: parallel-Jacobi marg, fused lin_vis_lm + DMMA asm_pairs + DMMA syrk: GPU tests, bench line, per-kernel probe of the variants, ncu of the new kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.log 2>&1
tail -4 gpurun_out/r02c_gpu_tests.log
grep -q passed gpurun_out/r02c_gpu_tests.log || { tail -60 gpurun_out/r02c_gpu_tests.log; }
PROBE_COPIES=16 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02c_probe_default.txt 2>&1
PROBE_COPIES=16 VIWB_SYRK_DFMA=1 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02c_probe_syrk_dfma.txt 2>&1
PROBE_COPIES=16 VIWB_NO_FUSED=1 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02c_probe_nofused.txt 2>&1
cat gpurun_out/r02c_probe_*.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -c 600 gpurun_out/r02c_bench.err
head -c 3000 gpurun_out/r02c_bench.json
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(marg|lin_vis_lm|asm_pairs|syrk_mma|solve|lin_small)_kernel' \
    -o gpurun_out/r02c_new python profiles/ncu_target.py --iters 1 --no-lk > gpurun_out/r02c_ncu.log 2>&1
VIWB_SYRK_DFMA=1 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^syrk_kernel' -c 1 \
    -o gpurun_out/r02c_syrk_dfma python profiles/ncu_target.py --iters 1 --no-lk >> gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log
xz -T0 -3 gpurun_out/r02c_new.ncu-rep gpurun_out/r02c_syrk_dfma.ncu-rep
ls -la gpurun_out/

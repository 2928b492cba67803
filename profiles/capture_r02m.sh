#!/bin/bash
# r02m: device marginalization behind the adapter; LK at 5 blocks/SM; per-config bench lines; ncu of the solver chain at this commit
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_gpu_tests.log 2>&1
tail -4 gpurun_out/r02m_gpu_tests.log
grep -q " passed" gpurun_out/r02m_gpu_tests.log || tail -60 gpurun_out/r02m_gpu_tests.log
timeout 200 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02m_lk_probe.txt 2>&1
cat gpurun_out/r02m_lk_probe.txt
PROBE_COPIES=16 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02m_probe.txt 2>&1
cat gpurun_out/r02m_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
tail -c 600 gpurun_out/r02m_bench.err
head -c 600 gpurun_out/r02m_bench.json
for c in 1 3 4 6; do
  timeout 600 python bench.py --steps 3 --warmup 3 --config $c --cpu-seconds 5 > gpurun_out/r02m_bench_c$c.json 2> gpurun_out/r02m_bench_c$c.err
  tail -c 400 gpurun_out/r02m_bench_c$c.err; head -c 500 gpurun_out/r02m_bench_c$c.json; echo
done
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(solve|lin_vis_lm|lin_small|marg_tri|marg_ql|marg_apply|lin_vis_lm_wide_marg)_kernel' \
    -o gpurun_out/r02m_chain python profiles/ncu_target.py --iters 1 --no-lk > gpurun_out/r02m_ncu.log 2>&1
tail -3 gpurun_out/r02m_ncu.log
xz -T0 -3 gpurun_out/r02m_chain.ncu-rep
ls -la gpurun_out/ | tail -20

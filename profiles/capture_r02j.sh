#!/bin/bash
# r02j: LK with UINT32-word tensor maps (TMA tiles), pipelined; occupancy variants; tests; bench; ncu of lk_track
set -x
mkdir -p gpurun_out
timeout 200 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so profiles/variants_lk6.so profiles/variants_lk5.so > gpurun_out/r02j_lk_probe.txt 2>&1
if ! grep -q "tick" gpurun_out/r02j_lk_probe.txt; then tail -5 gpurun_out/r02j_lk_probe.txt; echo "TMA PATH FAILED -- continuing without it"; export VIWB_LK_NO_TMA=1; fi
VIWB_LK_NO_TMA=1 timeout 300 python profiles/lk_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02j_lk_probe.txt 2>&1
cat gpurun_out/r02j_lk_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_gpu_tests.log 2>&1
tail -4 gpurun_out/r02j_gpu_tests.log
grep -q " passed" gpurun_out/r02j_gpu_tests.log || tail -60 gpurun_out/r02j_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
tail -c 600 gpurun_out/r02j_bench.err
head -c 1200 gpurun_out/r02j_bench.json
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(lk_track_tasks|marg_tri|marg_ql|marg_apply)_kernel' \
    -o gpurun_out/r02j_lk python profiles/ncu_target.py --iters 1 > gpurun_out/r02j_ncu.log 2>&1
tail -3 gpurun_out/r02j_ncu.log
xz -T0 -3 gpurun_out/r02j_lk.ncu-rep
ls -la gpurun_out/

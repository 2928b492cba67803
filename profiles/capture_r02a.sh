#!/bin/bash
# r02a: measured FP64 / DMMA / shared-memory denominators + one `ncu --set full` launch sequence of every hot kernel at HEAD (592 windows)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
./profiles/micro/microbench > gpurun_out/r02a_microbench.json 2> gpurun_out/r02a_microbench.err
cat gpurun_out/r02a_microbench.json
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(solve|marg|lin_vis|lm_reduce|lin_small|asm_items|syrk|lk_track_tasks|pyr_down_tasks)_kernel' \
    -o gpurun_out/r02a_all python profiles/ncu_target.py > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
ls -la gpurun_out/

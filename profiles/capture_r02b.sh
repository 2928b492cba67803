#!/bin/bash
# r02b: one `ncu --set full` launch sequence of every hot kernel at HEAD (592 windows, max_num_iterations 1), compressed to fit gpurun_out
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(solve|marg|lin_vis|lm_reduce|lm_reduce_wide|lin_small|asm_items|asm_items_split|syrk|lk_track_tasks|pyr_down_tasks)_kernel' \
    -o gpurun_out/r02b_all python profiles/ncu_target.py --iters 1 > gpurun_out/r02b_ncu.log 2>&1
tail -3 gpurun_out/r02b_ncu.log
ncu -i gpurun_out/r02b_all.ncu-rep --page raw --csv > gpurun_out/r02b_raw.csv 2>/dev/null
xz -T0 -3 gpurun_out/r02b_all.ncu-rep
ls -la gpurun_out/
du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then
  xz -dc gpurun_out/r02b_all.ncu-rep.xz > /tmp/all.ncu-rep
  ncu -i /tmp/all.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | xz -T0 -3 > gpurun_out/r02b_source.csv.xz
  rm gpurun_out/r02b_all.ncu-rep.xz
fi
ls -la gpurun_out/

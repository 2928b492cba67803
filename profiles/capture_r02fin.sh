#!/bin/bash
# r02fin: evidence at the final commit of the round (same command list as r02r, two commits later: LK border staging, landmark-block descriptors) -- ncu --set full of every kernel with >= 1 % of the step, the ncu launch list of the bench command, bench lines per configuration, reference arm
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02fin_gpu_tests.log 2>&1; tail -3 gpurun_out/r02fin_gpu_tests.log
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:'^(solve|lk_track_tasks|lin_vis_lm|lin_vis_lm_wide|lin_small|asm_pairs|asm_pairs_wide|pair_reduce|syrk_mma|marg_prep|marg_tri|marg_ql|marg_apply)_kernel' \
    -o gpurun_out/r02fin_all python profiles/ncu_target.py --iters 1 > gpurun_out/r02fin_ncu.log 2>&1
tail -3 gpurun_out/r02fin_ncu.log
xz -T0 -3 gpurun_out/r02fin_all.ncu-rep
ls -la gpurun_out/r02fin_all.ncu-rep.xz
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02fin_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-profile --cpu-seconds 1 --e2e-lanes 1 --parity-windows 16 > gpurun_out/r02fin_bench_under_ncu.log 2>&1
wc -l gpurun_out/r02fin_launches.csv
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02fin_bench.json 2> gpurun_out/r02fin_bench.err
tail -c 400 gpurun_out/r02fin_bench.err
head -c 300 gpurun_out/r02fin_bench.json; echo
for c in 1 3 4 6; do
  timeout 600 python bench.py --steps 5 --warmup 3 --config $c > gpurun_out/r02fin_bench_c$c.json 2> gpurun_out/r02fin_bench_c$c.err
  tail -c 300 gpurun_out/r02fin_bench_c$c.err; head -c 300 gpurun_out/r02fin_bench_c$c.json; echo
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r02fin_bench_reference.json 2> gpurun_out/r02fin_bench_reference.err
head -c 300 gpurun_out/r02fin_bench_reference.json; echo
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02fin_smoke.log 2>&1
tail -2 gpurun_out/r02fin_smoke.log
ls -la gpurun_out | tail -20

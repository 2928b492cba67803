#!/bin/bash
# r02n: lin_vis_lm reduction phases restructured, pair_reduce item staging, transposed prior Jacobian, parallel fetch unpack; occupancy / block-size variants
set -x
mkdir -p gpurun_out
PROBE_COPIES=48 timeout 900 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so profiles/var_lvl5.so profiles/var_lsm3.so profiles/var_lsm4.so profiles/var_tri128.so profiles/var_prn128.so profiles/var_prn320.so > gpurun_out/r02n_probe.txt 2>&1
cat gpurun_out/r02n_probe.txt
PROBE_COPIES=48 PROBE_CONFIG=4 timeout 900 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so profiles/var_lvl5.so >> gpurun_out/r02n_probe.txt 2>&1
tail -2 gpurun_out/r02n_probe.txt
VIWB_TIMING=1 timeout 300 python profiles/e2e_probe.py 2> gpurun_out/r02n_e2e_probe.txt
grep -v "^$" gpurun_out/r02n_e2e_probe.txt | tail -40
timeout 900 python bench.py --steps 5 --warmup 3 --e2e-lanes-sweep 2,6,8,12 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
tail -c 1500 gpurun_out/r02n_bench.err
head -c 400 gpurun_out/r02n_bench.json

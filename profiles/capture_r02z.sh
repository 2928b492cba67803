#!/bin/bash
# r02z: e2e lanes with the camera-tick upload arguments marshalled once
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --e2e-lanes-sweep 6 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
tail -c 600 gpurun_out/r02z_bench.err
head -c 300 gpurun_out/r02z_bench.json
timeout 600 python bench.py --steps 5 --warmup 3 --config 6 > gpurun_out/r02z_bench_c6.json 2> gpurun_out/r02z_bench_c6.err
head -c 300 gpurun_out/r02z_bench_c6.json

#!/bin/bash
# r02h: why does every TMA tensor load raise "illegal instruction" here?  environment + isolation variants; then the marg three-kernel pipeline and LK occupancy
set -x
mkdir -p gpurun_out
{ echo "LD_PRELOAD=$LD_PRELOAD"; env | grep -i -E "cuda|nvidia|nvbit|inject|cupti|sanitizer" ; nvidia-smi -q | grep -i -E "mig mode|virtualization|confidential|product name|driver version|cuda version" ; } > gpurun_out/r02h_env.txt 2>&1
cat gpurun_out/r02h_env.txt
for v in "5" "4" "1 0 96" "1 0 94" "2 1 96"; do timeout 60 ./profiles/micro/tma_test $v; done > gpurun_out/r02h_tma_test.txt 2>&1
for v in "4" "1 0 96"; do env -u LD_PRELOAD -u CUDA_INJECTION64_PATH -u NVBIT_TOOL timeout 60 ./profiles/micro/tma_test $v; done >> gpurun_out/r02h_tma_test.txt 2>&1
cat gpurun_out/r02h_tma_test.txt
compute-sanitizer --tool memcheck ./profiles/micro/tma_test 1 0 96 > gpurun_out/r02h_sanitizer.txt 2>&1; tail -25 gpurun_out/r02h_sanitizer.txt
export VIWB_LK_NO_TMA=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_gpu_tests.log 2>&1
tail -4 gpurun_out/r02h_gpu_tests.log
grep -q " passed" gpurun_out/r02h_gpu_tests.log || tail -60 gpurun_out/r02h_gpu_tests.log
PROBE_COPIES=16 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so > gpurun_out/r02h_probe.txt 2>&1
PROBE_COPIES=16 VIWB_MARG_ONE_KERNEL=1 timeout 600 python profiles/kernel_probe.py viw-fusion_b200/csrc/libviwb.so >> gpurun_out/r02h_probe.txt 2>&1
cat gpurun_out/r02h_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
tail -c 600 gpurun_out/r02h_bench.err
head -c 1200 gpurun_out/r02h_bench.json
ls -la gpurun_out/

#!/bin/bash
# r02i: which tensor-map shapes load on this box (tma_probe matrix); host-side phase timing of the host-buffer call
set -x
mkdir -p gpurun_out
for v in "0 32 32 0 0 96" "0 32 32 0 0 752" "0 64 32 0 0 752" "0 128 8 0 0 752" "0 32 32 1 0 752" "0 32 32 0 2 752" "0 16 16 0 0 752" "2 8 32 0 0 96" "2 16 16 0 0 752" "2 32 8 3 0 752" "7 8 32 0 0 96" "7 32 8 3 0 752" "7 32 8 3 2 752" "6 64 8 3 2 752"; do timeout 60 ./profiles/micro/tma_probe $v; done > gpurun_out/r02i_tma_probe.txt 2>&1
cat gpurun_out/r02i_tma_probe.txt
python - <<'P' > gpurun_out/r02i_torch_tma.txt 2>&1
import torch
a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
b = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
c = a @ b
torch.cuda.synchronize()
print("torch bf16 matmul ok", float(c.float().abs().mean()))
P
cat gpurun_out/r02i_torch_tma.txt
VIWB_LK_NO_TMA=1 VIWB_TIMING=1 timeout 300 python profiles/e2e_probe.py > gpurun_out/r02i_e2e_probe.txt 2>&1
tail -12 gpurun_out/r02i_e2e_probe.txt

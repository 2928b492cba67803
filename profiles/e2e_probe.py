import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/viw-fusion_b200/python')
import bench
from viwb import abi, lib
ctx = lib.Context(0)
cfg, seqs, first = bench.make_windows(0, 16, 1)
a0,_,q0 = ctx.optimization_batch([f[0] for f in first],[f[1] for f in first],[abi.MARGIN_OLD]*len(first))
probs, states = bench.replicate(seqs, q0, a0, 64, 0)
call = ctx.prepare_optimization_batch(probs, states, [abi.MARGIN_OLD]*len(probs))
for i in range(3):
    t=time.perf_counter(); call(); print('call %.1f ms'%((time.perf_counter()-t)*1e3), file=sys.stderr)

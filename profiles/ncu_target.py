"""ncu target at the bench shape: B windows of the bench's C2 workload (+ one stereo camera tick of LK), warm-up passes outside the
profiled region, then ONE pass between cudaProfilerStart / cudaProfilerStop with max_num_iterations = 2 (round 1 is a typical
iteration: accept decision + assembly + Schur + Cholesky + dogleg), so that `ncu --profile-from-start off` captures one launch
sequence of every kernel instead of nine.
Usage (GPU box): ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'...' -o out python profiles/ncu_target.py
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
import bench  # noqa: E402
from viwb import abi, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--distinct", type=int, default=8)
ap.add_argument("--copies", type=int, default=74)       # 8 x 74 = 592 windows = 4 x 148 SMs
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--no-lk", action="store_true")
a = ap.parse_args()
rt = None
for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
ctx = lib.Context(0)
cfg, seqs, first = bench.make_windows(0, a.distinct, a.copies, a.config)
a0, _, q0 = ctx.optimization_batch([f[0] for f in first], [f[1] for f in first], [abi.MARGIN_OLD] * len(first))
probs, states = bench.replicate(seqs, q0, a0, a.copies, 0)
B = len(probs)
opt = abi.default_options() if hasattr(abi, "default_options") else None
if opt is not None:
    opt.max_num_iterations = a.iters
batch = ctx.batch(probs, states, [abi.MARGIN_OLD] * B, opt)
lk = None
if not a.no_lk:
    scenes = bench.make_scenes(0, 4)
    feed = bench.FrameFeed(ctx, scenes, B)
    lk = ctx.lk_batch(B, bench.IMG_W, bench.IMG_H, bench.N_FEAT, stereo=True, flow_back=True)
    lk.upload(prev=feed.left[0], cur=feed.left[1], right=feed.right[1], prev_pts=feed.pts[0], n_prev=feed.n, stereo_pts=feed.pts[1], n_stereo=feed.n)
    lk.run(); lk.download()
for _ in range(2):
    batch.run()
    if lk is not None:
        lk.run()
batch.download()
if rt is not None:
    rt.cudaProfilerStart()
batch.run()
if lk is not None:
    lk.run()
sts, sums, pri = batch.download()
if rt is not None:
    rt.cudaProfilerStop()
print("ok", B, sums[0].num_iterations, ctx.launch_count())

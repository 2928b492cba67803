"""Small driver for ncu captures: one batch of C2 steady-state windows, `--runs` passes of the hot path.
Usage (on the GPU box):  ncu ... python profiles/run_profile.py --distinct 8 --copies 8 --runs 2"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
import bench  # noqa: E402
from viwb import abi, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--distinct", type=int, default=8)
ap.add_argument("--copies", type=int, default=8)
ap.add_argument("--runs", type=int, default=2)
a = ap.parse_args()
ctx = lib.Context(0)
cfg, seqs, first = bench.make_windows(0, a.distinct, a.copies)
a0, _, q0 = ctx.optimization_batch([f[0] for f in first], [f[1] for f in first], [abi.MARGIN_OLD] * len(first))
probs, states = bench.replicate(seqs, q0, a0, a.copies, 0)
batch = ctx.batch(probs, states, [abi.MARGIN_OLD] * len(probs))
for _ in range(a.runs):
    batch.run()
sts, sums, pri = batch.download()
print("ok", len(probs), sums[0].num_iterations, ctx.launch_count())

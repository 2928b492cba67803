"""Host-side breakdown of the host-buffer entry point (VIWB_TIMING=1 prints build / execute / fetch per call on stderr): one window, 16 windows, 1024 windows.
Usage (GPU box): VIWB_TIMING=1 python profiles/e2e_probe.py 2> out.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
import bench  # noqa: E402
from viwb import abi, lib  # noqa: E402

ctx = lib.Context(0)
cfg, seqs, first = bench.make_windows(0, 16, 1)
a0, _, q0 = ctx.optimization_batch([f[0] for f in first], [f[1] for f in first], [abi.MARGIN_OLD] * len(first))
for copies in (0, 1, 64):
    probs, states = bench.replicate(seqs, q0, a0, max(copies, 1), 0)
    if copies == 0:
        probs, states = probs[:1], states[:1]
    call = ctx.prepare_optimization_batch(probs, states, [abi.MARGIN_OLD] * len(probs))
    for i in range(4):
        t = time.perf_counter()
        call()
        print("B=%d call %.2f ms" % (len(probs), (time.perf_counter() - t) * 1e3), file=sys.stderr)

"""Per-kernel CUDA-event times of one optimisation batch for alternative builds of the library.
Usage (GPU box): python profiles/kernel_probe.py libA.so libB.so ...   (env PROBE_DISTINCT, PROBE_COPIES)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
import bench  # noqa: E402
from viwb import abi, lib  # noqa: E402

D, R = int(os.environ.get("PROBE_DISTINCT", "37")), int(os.environ.get("PROBE_COPIES", "32"))
cfg, seqs, first = bench.make_windows(0, D, R, int(os.environ.get("PROBE_CONFIG", "2")))
for path in sys.argv[1:]:
    ctx = lib.Context(0, os.path.abspath(path))
    a0, _, q0 = ctx.optimization_batch([f[0] for f in first], [f[1] for f in first], [abi.MARGIN_OLD] * len(first))
    probs, states = bench.replicate(seqs, q0, a0, R, 0)
    batch = ctx.batch(probs, states, [abi.MARGIN_OLD] * len(probs))
    for _ in range(2):
        batch.run()
    batch.download()
    ctx.set_profiling(True)
    for _ in range(3):
        batch.run()
    sts, sums, pri = batch.download()
    prof = ctx.profile()
    ctx.set_profiling(False)
    tot = sum(v[0] for v in prof.values()) / 3
    print(os.path.basename(path), "B=%d step %.2f ms" % (len(probs), tot), {k: round(v[0] / max(1, v[1]), 3) for k, v in prof.items()},
          "iters", sums[0].num_iterations)
    batch.destroy()
    ctx.close()

"""Worker of tests/test_multi_rank.py: one process per 'GPU' (here: per CPU rank, gloo), exactly the plumbing bench.py uses for
--gpus N -- rank-sharded sequences, no data-path collective, barrier, max-over-ranks timing, whole-job aggregate -- with the
kernel-logic emulation standing in for the device."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from viwb import abi, lib  # noqa: E402
from emu import build_emu  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
ctx = lib.Context(0, build_emu.OUT)          # built by the parent test before the ranks start
cfg, seqs, first = bench.make_windows(rank, 2, 1)
a0, _, q0 = ctx.optimization_batch([f[0] for f in first], [f[1] for f in first], [abi.MARGIN_OLD] * 2)
probs, states = bench.replicate(seqs, q0, a0, 2, rank)
sts, sums, pri = ctx.optimization_batch(probs, states, [abi.MARGIN_OLD] * len(probs))
dist.barrier()
fake_ms = 10.0 * (rank + 1)                          # rank 1 is the slow replica
ms = bench.rank_max(fake_ms, world, device="cpu")
value = bench.job_throughput(world, len(probs), 1, ms * 1e-3)
sig = float(np.sum([s[:77].sum() for s in sts]))     # depends on the rank's own sequences
gathered = [None] * world
dist.all_gather_object(gathered, {"rank": rank, "sig": sig, "n": len(probs), "iters": [int(s.num_iterations) for s in sums]})
if rank == 0:
    json.dump({"ms": ms, "value": value, "ranks": gathered}, open(sys.argv[1], "w"))
dist.destroy_process_group()
ctx.close()

"""N>1 path of bench.py on CPU: world_size 2, gloo, launched the way the driver launches the bench (torch.distributed.run,
127.0.0.1 rendezvous).  The window solve does not shard (replicas only), so what is checked is the replica
plumbing: rank-distinct sequences, barrier, max-over-ranks time, whole-job aggregate."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_replicas_gloo(tmp_path):
    from emu import build_emu
    build_emu.build()
    out = str(tmp_path / "ranks.json")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "multi_rank_worker.py"), out]
    subprocess.run(cmd, check=True, env=env, timeout=600, cwd=ROOT)
    r = json.load(open(out))
    assert r["ms"] == 20.0                                   # the slow rank sets the time
    assert abs(r["value"] - 2 * 4 * 1 / 0.020) < 1e-9        # whole-job aggregate over both ranks
    ranks = sorted(r["ranks"], key=lambda d: d["rank"])
    assert [d["rank"] for d in ranks] == [0, 1] and all(d["n"] == 4 for d in ranks)
    assert ranks[0]["sig"] != ranks[1]["sig"]                # each rank worked on its own sequences
    assert all(2 <= it <= 9 for d in ranks for it in d["iters"])

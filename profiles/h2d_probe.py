"""Where the end-to-end step's upload time goes: the stereo camera tick's images (page-locked numpy arrays -> viwb_lk_batch_upload) and the window tables
(caller tables -> pinned slab -> device), each alone and both from two threads.  Usage (GPU box): python profiles/h2d_probe.py"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
import torch  # noqa: E402
import bench  # noqa: E402
from viwb import abi, lib  # noqa: E402

B = 1776
ctx = lib.Context(0)
scenes = bench.make_scenes(0, 8)
feed = bench.FrameFeed(ctx, scenes, B, True)
lk = ctx.lk_batch(B, bench.IMG_W, bench.IMG_H, bench.N_FEAT, stereo=True, flow_back=True)
lk.upload(**feed.tick_args(1, first=True))
lk.run()
lk.download()
img_bytes = feed.tick_bytes()


def t_images(n=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        lk.upload(**feed.tick_args(i % 2))
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


dt = t_images()
print("images alone: %.2f ms per tick, %.1f MB -> %.1f GB/s" % (dt * 1e3, img_bytes / 1e6, img_bytes / dt / 1e9))

cfg, seqs, first = bench.make_windows(0, 37, 48, 2)
ctx2 = lib.Context(0)
a0, _, q0 = ctx2.optimization_batch([f[0] for f in first], [f[1] for f in first], [abi.MARGIN_OLD] * len(first))
probs, states = bench.replicate(seqs, q0, a0, 48, 0)


def t_tables(n=3):
    torch.cuda.synchronize()
    u0 = ctx2.h2d_bytes()
    t0 = time.perf_counter()
    for _ in range(n):
        b = ctx2.batch(probs, states, [abi.MARGIN_OLD] * B)
        torch.cuda.synchronize()
        b.destroy()
    dt = (time.perf_counter() - t0) / n
    return dt, (ctx2.h2d_bytes() - u0) / n


dt, nb = t_tables()
print("tables alone (lowering + upload, no kernels): %.2f ms per batch, %.1f MB on the wire" % (dt * 1e3, nb / 1e6))

res = {}
th = [threading.Thread(target=lambda: res.__setitem__("img", t_images(6))), threading.Thread(target=lambda: res.__setitem__("tab", t_tables(3)))]
t0 = time.perf_counter()
for t in th:
    t.start()
for t in th:
    t.join()
print("both at once: images %.2f ms per tick, tables %.2f ms per batch" % (res["img"] * 1e3, res["tab"][0] * 1e3))

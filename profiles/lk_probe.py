"""Times one batched LK tick (viwb_lk_batch_run) for alternative builds of the library.
Usage (GPU box): python profiles/lk_probe.py libA.so libB.so ...   -> ms per tick per build"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
import bench  # noqa: E402
from viwb import lib  # noqa: E402

F = int(os.environ.get("LK_STREAMS", "512"))
scenes = bench.make_scenes(0, 4)
for path in sys.argv[1:]:
    ctx = lib.Context(0, os.path.abspath(path))
    feed = bench.FrameFeed(None, scenes, F)
    lk = ctx.lk_batch(F, bench.IMG_W, bench.IMG_H, bench.N_FEAT, stereo=True, flow_back=True)
    lk.upload(prev=feed.left[0], cur=feed.left[1], right=feed.right[1], prev_pts=feed.pts[0], n_prev=feed.n, stereo_pts=feed.pts[1], n_stereo=feed.n)
    for _ in range(3):
        lk.run()
    lk.download()
    ctx.set_profiling(True)
    t0 = time.perf_counter()
    for _ in range(5):
        lk.run()
    out = lk.download()
    dt = (time.perf_counter() - t0) / 5
    prof = ctx.profile()
    ctx.set_profiling(False)
    print(os.path.basename(path), "tick %.3f ms" % (dt * 1e3), {k: round(v[0] / max(1, v[1]), 3) for k, v in prof.items()}, "tracked %.3f" % out[1].mean())
    lk.close()
    ctx.close()

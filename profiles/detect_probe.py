"""Times one batched detector tick (setMask + goodFeaturesToTrack on the tracker's resident images) per build, and
cv2.goodFeaturesToTrack on the host for the same frames.
Usage (GPU box): python profiles/detect_probe.py libA.so [libB.so ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
import bench  # noqa: E402
from viwb import lib  # noqa: E402

F = int(os.environ.get("DET_STREAMS", "296"))
scenes = bench.make_scenes(0, 4)
rng = np.random.default_rng(0)
for path in sys.argv[1:]:
    ctx = lib.Context(0, os.path.abspath(path))
    feed = bench.FrameFeed(None, scenes, F)
    lk = ctx.lk_batch(F, bench.IMG_W, bench.IMG_H, 192, stereo=False, flow_back=True)
    n = np.full(F, 110, np.int32)                                    # 110 tracked points survive, 40 corners to find
    pts = np.zeros((F, 192, 2), np.float32); pts[:, :150] = feed.pts[1][:, :150]
    cnt = np.zeros((F, 192), np.int32); cnt[:, :150] = rng.integers(1, 12, (F, 150))
    lk.upload(prev=feed.left[0], cur=feed.left[1], prev_pts=pts, n_prev=n)
    lk.run(); lk.download()
    det = ctx.detector(F, bench.IMG_W, bench.IMG_H, 192, 30)
    for _ in range(2):
        det.detect(None, pts, cnt, n, 150, resident=lk)
    ctx.set_profiling(True)
    t0 = time.perf_counter()
    for _ in range(5):
        keep, n_keep, new_pts, n_new, _ = det.detect(None, pts, cnt, n, 150, resident=lk)
    dt = (time.perf_counter() - t0) / 5
    prof = ctx.profile()
    ctx.set_profiling(False)
    per = {k: round(v[0] / max(1, v[1]), 3) for k, v in prof.items()}
    dev_ms = sum(per.values())
    print(os.path.basename(path), "F=%d tick %.3f ms wall, %.3f ms device" % (F, dt * 1e3, dev_ms), per, "kept %.1f new %.1f" % (n_keep.mean(), n_new.mean()),
          "%.0f frames/s device, image bytes %.1f GB/s" % (F / dev_ms * 1e3, det.algorithmic_bytes() / dev_ms * 1e-6))
    det.close(); lk.close(); ctx.close()
import cv2  # noqa: E402
imgs = feed.left[1][:16]
t0 = time.perf_counter()
for f in range(len(imgs)):
    mask = np.full(imgs[f].shape, 255, np.uint8)
    for p in pts[f, :110]:
        cv2.circle(mask, (int(round(p[0])), int(round(p[1]))), 30, 0, -1)
    cv2.goodFeaturesToTrack(imgs[f], 40, 0.01, 30, mask=mask)
print("cv2 setMask+goodFeaturesToTrack: %.3f ms per frame (1 host thread pool = %d)" % ((time.perf_counter() - t0) / len(imgs) * 1e3, cv2.getNumThreads()))

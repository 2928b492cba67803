"""Session tracker (viwb_tracker_*: one FeatureTracker::trackImage() per stream per tick, state resident in HBM) on F streams:
end-to-end ticks (host images in, featureFrame rows out), the per-kernel CUDA-event times of a tick, and the same trackImage()
restated over cv2 (tests/parity_checks.py:time_tracker_reference -> the checker's FeatureTrackerRef, real cv2.calcOpticalFlowPyrLK /
cv2.goodFeaturesToTrack) on one host thread.  Writes one JSON line.  Usage (GPU box): python profiles/track_probe.py [libviwb.so]   (TRK_STREAMS, TRK_TICKS env)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from viwb import lib  # noqa: E402
import parity_checks as pc  # noqa: E402

F = int(os.environ.get("TRK_STREAMS", "296"))
TICKS = int(os.environ.get("TRK_TICKS", "8"))
W, H, MAX_CNT, MIN_DIST = 752, 480, 150, 30
CAM0 = (461.1586, 459.7529, 362.6593, 248.5236, -0.2847798, 0.08245052, -1.0946e-06, 4.78701e-06)     # config/euroc/cam0_pinhole.yaml
CAM1 = (457.5874, 456.1340, 379.9994, 255.2381, -0.2836831, 0.07395907, 1.9359e-04, 1.7618e-05)       # config/euroc/cam1_pinhole.yaml

DISTINCT = 4
seqs = [pc.camera_sequence(300 + k, W, H, TICKS + 2) for k in range(DISTINCT)]
idx = np.arange(F) % DISTINCT
left = [np.ascontiguousarray(np.stack([seqs[i][0][t] for i in idx])) for t in range(TICKS + 2)]
right = [np.ascontiguousarray(np.stack([seqs[i][1][t] for i in idx])) for t in range(TICKS + 2)]

ctx = lib.Context(0, os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else None)
for a in left + right:
    ctx.host_register(a)
trk = ctx.tracker(F, W, H, CAM0, CAM1, MAX_CNT, MIN_DIST, True)
for t in range(2):                                   # warm-up ticks: detection of the first corners, first temporal flow
    trk.track(0.05 * (t + 1), left[t], right[t]); trk.download()
l0 = ctx.launch_count()
t0 = time.perf_counter()
rows = []
for t in range(2, TICKS + 2):
    trk.track(0.05 * (t + 1), left[t], right[t])
    out = trk.download()
    rows.append((float(out[0].mean()), float(out[4].mean()), float((out[2][:, :MAX_CNT] > 1).sum() / F)))
wall = (time.perf_counter() - t0) / TICKS
launches = (ctx.launch_count() - l0) / TICKS
# per-kernel pass (CUDA events around every launch), two more ticks on the same session
ctx.set_profiling(True)
for t in range(2):
    trk.track(0.05 * (TICKS + 3 + t), left[t], right[t]); trk.download()
prof = ctx.profile()
ctx.set_profiling(False)
per = {k: {"ms_per_launch": round(v[0] / max(1, v[1]), 4), "launches_per_tick": v[1] / 2} for k, v in prof.items()}
dev_ms = max(sum(v[0] for v in prof.values()) / 2, 1e-9)
alg = trk.algorithmic_bytes()
trk.close()

# CPU: the restated trackImage() over cv2, one stream, one thread (the checker lives under tests/ + oracle/; this script only reports its time)
cpu_ms = pc.time_tracker_reference(CAM0, CAM1, MAX_CNT, MIN_DIST, seqs[0], TICKS)

print(json.dumps({"what": "viwb_tracker_track + download, stereo 752x480, MAX_CNT 150, MIN_DIST 30, FLOW_BACK 1", "streams": F, "ticks": TICKS,
                  "e2e_ms_per_tick": wall * 1e3, "e2e_frames_per_s": F / wall, "device_ms_per_tick": dev_ms, "device_frames_per_s": F / dev_ms * 1e3,
                  "h2d_bytes_per_tick": int(2 * F * W * H), "launches_per_tick": launches, "kernels": per,
                  "algorithmic_bytes_per_tick": alg, "algorithmic_gbs_device": alg / dev_ms * 1e-6,
                  "rows_left_right_tracked_mean_last": rows[-1],
                  "cpu_trackImage_ms_per_frame_1thread": cpu_ms, "cpu_frames_per_s_1thread": 1e3 / cpu_ms}))

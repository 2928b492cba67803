"""Runs a cross-section of the parity checks against the sanitizer build of the kernel-logic emulation (tests/emu/build_emu.py:build_sanitized).
Started by tests/test_emu_sanitized.py in a child process with libasan preloaded; prints one `ok <name>` line per check."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "viw-fusion_b200", "python")]
from viwb import lib as L            # noqa: E402
import viw_oracle as oracle          # noqa: E402
import parity_checks as pc           # noqa: E402

emu = L.Context(0, sys.argv[1])
CHECKS = [
    ("alignment", lambda: [pc.check_visual_imu_alignment(emu, oracle, c) for c in (1, 4)]),
    ("tracker_edges", lambda: pc.check_tracker_edges(emu)),
    ("tracker_session", lambda: pc.check_tracker_session(emu, streams=2, w=240, h=180, ticks=3, max_cnt=40, min_dist=18)),
    ("edge_cases", lambda: pc.check_edge_cases(emu, oracle)),
    ("small_edges", lambda: pc.check_small_edges(emu, oracle)),
    ("solve_prior_chain", lambda: pc.check_solve(emu, oracle, 4, prior_chain=True)),
    ("solve_stereo", lambda: pc.check_solve(emu, oracle, 2)),
    ("marginalize", lambda: pc.check_marginalize(emu, oracle, 4)),
    ("batch", lambda: pc.check_batch_matches_single(emu, oracle)),
    ("preintegration", lambda: pc.check_preintegration(emu, oracle)),
    ("outliers", lambda: pc.check_outlier_rejection(emu, oracle)),
    ("triangulation", lambda: pc.check_triangulation(emu, oracle)),
    ("undistort", lambda: pc.check_undistort_velocity(emu)),
    ("lk", lambda: pc.check_lk(emu)),
    ("lk_batch", lambda: pc.check_lk_batch(emu, streams=2, w=200, h=160, min_both=8)),
    ("set_mask", lambda: pc.check_set_mask(emu)),
    ("detector", lambda: pc.check_detector_batch(emu, streams=2)),
    ("reanchor", lambda: pc.check_reanchor(emu, oracle, 4)),
]
for name, fn in CHECKS:
    fn()
    print("ok", name, flush=True)
print("checks", len(CHECKS), flush=True)

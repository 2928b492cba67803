"""Generates the committed fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

lk_golden.npz      OUTPUTS OF THE REFERENCE'S OWN TRACKER: cv2.calcOpticalFlowPyrLK (OpenCV, the third-party library
                   FeatureTracker::trackImage calls) on a small synthetic image pair, for the three call shapes of
                   feature_tracker.cpp (:139 maxLevel 3; :125/:145 maxLevel 1 + OPTFLOW_USE_INITIAL_FLOW; maxLevel 0).
tracker_golden.npz FeatureTracker::trackImage() over a 5-tick synthetic stereo sequence (160 x 120): the line-by-line restatement
                   oracle/feature_oracle.py:FeatureTrackerRef with cv2.calcOpticalFlowPyrLK AND cv2.goodFeaturesToTrack (scalar path)
                   underneath, i.e. the reference's own third-party calls: ids, track counts, pixel / undistorted points, velocities
                   of both cameras per tick, plus the images.
window_golden.npz  Regression vectors of the CPU oracle (oracle/; its Ceres-solve half is PARITY UNPINNED: the reference ships no known-answer
                   vectors for the window solve and ceres-solver is not in the image) for two seeded windows: solved state, iteration
                   count, costs and the marginalisation prior's information form (J^T J, J^T r).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))


def lk_fixture():
    import cv2
    import parity_checks as pc
    img0, img1, pts = pc.lk_images(7, 200, 160)
    h, w = img0.shape
    extra = np.array([[3.0, 4.0], [w - 2.5, h - 3.0], [w / 2, 1.0], [0.2, h / 2]], np.float32)
    pts = np.vstack([pts, extra]).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    out = {"img0": img0, "img1": img1, "pts": pts, "opencv_version": np.array(cv2.__version__)}
    for name, ml, flags in (("L3", 3, 0), ("L1_init", 1, cv2.OPTFLOW_USE_INITIAL_FLOW), ("L0", 0, 0)):
        init = (pts + np.float32([1.5, -1.0])) if flags else None
        p, st, err = cv2.calcOpticalFlowPyrLK(img0, img1, pts.reshape(-1, 1, 2), None if init is None else init.reshape(-1, 1, 2).copy(),
                                              winSize=(21, 21), maxLevel=ml, criteria=crit, flags=flags)
        out[name + "_next"] = p.reshape(-1, 2)
        out[name + "_status"] = st.reshape(-1)
        out[name + "_err"] = err.reshape(-1)
        if init is not None:
            out[name + "_init"] = init
    np.savez_compressed(os.path.join(HERE, "lk_golden.npz"), **out)


TRK = dict(w=160, h=120, ticks=5, max_cnt=30, min_dist=14, seed=91,
           cam0=(98.1, 97.8, 78.2, 61.3, -0.2847798, 0.08245052, -1.0946e-06, 4.78701e-06),
           cam1=(97.4, 97.1, 82.9, 58.8, -0.2836831, 0.07395907, 1.9359e-04, 1.7618e-05))


def tracker_fixture():
    import cv2
    import feature_oracle as fo
    import parity_checks as pc
    was = cv2.useOptimized()
    cv2.setUseOptimized(False)             # goodFeaturesToTrack's scalar path is the specification (tests/test_feature_oracle.py)
    try:
        left, right, _ = pc.camera_sequence(TRK["seed"], TRK["w"], TRK["h"], TRK["ticks"], disparity=4.0)
        ref = fo.FeatureTrackerRef(TRK["cam0"], TRK["cam1"], TRK["max_cnt"], TRK["min_dist"], True, use_cv_detector=True)
        out = {"left": np.stack(left), "right": np.stack(right), "opencv_version": np.array(cv2.__version__)}
        for t in range(TRK["ticks"]):
            r = ref.track_image(0.05 * (t + 1), left[t], right[t])
            for name, a in zip(("ids", "cnt", "pts", "un", "vel", "ids_r", "pts_r", "un_r", "vel_r"), r):
                out["t%d_%s" % (t, name)] = np.asarray(a)
    finally:
        cv2.setUseOptimized(was)
    np.savez_compressed(os.path.join(HERE, "tracker_golden.npz"), **out)


def window_fixture():
    import viw_oracle as vo
    from viwb import abi, synth
    out = {}
    for cid in (1, 4):
        cfg = synth.make_config(cid)
        seq = synth.Sequence(cfg, 3, 12)
        prob, st, _ = seq.window(0)
        a, sm, q = vo.optimization(prob, st, abi.MARGIN_OLD)
        A, b = q.information()
        k = "C%d_" % cid
        out[k + "x0"] = st
        out[k + "x"] = a
        out[k + "iters"] = np.array([sm.num_iterations, sm.num_successful_steps, sm.termination_type])
        out[k + "cost"] = np.array([sm.initial_cost, sm.final_cost])
        out[k + "prior_n"] = np.array([q.n])
        out[k + "prior_A"] = A
        out[k + "prior_b"] = b
    np.savez_compressed(os.path.join(HERE, "window_golden.npz"), **out)


if __name__ == "__main__":
    lk_fixture()
    tracker_fixture()
    window_fixture()
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])

"""Pins oracle/feature_oracle.py (setMask + goodFeaturesToTrack restatement) against the cv2 wheel of this image.
cv2's scalar code path (setUseOptimized(False)) is the specification and must match bit for bit; the SIMD build is
allowed its one-ulp eigenvalue differences but has to pick the same corners on these images."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import feature_oracle as fo
from viwb import synth


@pytest.fixture()
def scalar_cv():
    was = cv2.useOptimized()
    cv2.setUseOptimized(False)
    yield
    cv2.setUseOptimized(was)


def random_mask(rng, h, w, n, r=30):
    mask = np.full((h, w), 255, np.uint8)
    for _ in range(n):
        cv2.circle(mask, (int(rng.integers(0, w)), int(rng.integers(0, h))), r, 0, -1)
    return mask


@pytest.mark.parametrize("r", [1, 2, 3, 7, 30, 31])
def test_circle_matches_cv(r):
    hw = fo.circle_half_widths(r)
    for (cx, cy) in [(40, 40), (2, 77), (79, 3), (0, 0)]:
        a = np.full((80, 80), 255, np.uint8); b = a.copy()
        cv2.circle(a, (cx, cy), r, 0, -1)
        fo.paint_circle(b, cx, cy, r, hw)
        assert np.array_equal(a, b)


def test_min_eigen_map_bit_exact(scalar_cv):
    for seed, (h, w) in enumerate([(480, 752), (120, 161), (33, 47)]):
        img = synth.texture_image(h, w, seed)
        assert np.array_equal(cv2.cornerMinEigenVal(img, 3, ksize=3), fo.corner_min_eigen_val(img))


@pytest.mark.parametrize("seed", range(3))
def test_gftt_matches_cv(scalar_cv, seed):
    rng = np.random.default_rng(seed)
    img = synth.texture_image(480, 752, seed)
    mask = random_mask(rng, 480, 752, 100)
    for mc, use_mask, md in [(150, True, 30), (40, True, 30), (0, True, 30), (150, False, 30), (500, False, 7.5), (60, True, 0.5)]:
        ref = cv2.goodFeaturesToTrack(img, mc, 0.01, md, mask=mask if use_mask else None)
        ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
        got = fo.good_features_to_track(img, mc, 0.01, md, mask if use_mask else None)
        assert ref.shape == got.shape and np.array_equal(ref, got)


def test_gftt_same_corners_as_simd_build():
    was = cv2.useOptimized(); cv2.setUseOptimized(True)
    try:
        img = synth.texture_image(480, 752, 5)
        ref = cv2.goodFeaturesToTrack(img, 150, 0.01, 30).reshape(-1, 2)
    finally:
        cv2.setUseOptimized(was)
    got = fo.good_features_to_track(img, 150, 0.01, 30.0)
    common = {tuple(p) for p in ref} & {tuple(p) for p in got}
    assert len(common) >= 0.99 * len(ref)


def test_set_mask_follows_reference_loop():
    rng = np.random.default_rng(3)
    h, w, n = 480, 752, 180
    pts = np.stack([rng.uniform(1, w - 2, n), rng.uniform(1, h - 2, n)], 1).astype(np.float32)
    cnt = rng.integers(1, 12, n)
    mask, keep = fo.set_mask(w, h, pts, cnt, 30)
    # the loop of feature_tracker.cpp:78-88 with cv2 doing the drawing, same (stable) visiting order
    ref = np.full((h, w), 255, np.uint8); kept = []
    for i in np.argsort(-cnt, kind="stable"):
        c = (int(np.rint(pts[i, 0])), int(np.rint(pts[i, 1])))
        if ref[c[1], c[0]] == 255:
            kept.append(i); cv2.circle(ref, c, 30, 0, -1)
    assert np.array_equal(mask, ref) and list(keep) == kept
    assert all(cnt[keep[i]] >= cnt[keep[i + 1]] for i in range(len(keep) - 1))

"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the feature-detection half of FeatureTracker::trackImage().

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product never does.

What is restated
  * FeatureTracker::setMask()                       vins_estimator/src/featureTracker/feature_tracker.cpp:59-89
  * cv::goodFeaturesToTrack(cur_img, n_pts, MAX_CNT - n, 0.01, MIN_DIST, mask)        feature_tracker.cpp:192

Both lean on OpenCV, a third-party dependency that is not under /root/reference (system OpenCV 3.x of ROS Kinetic / Melodic, version
unpinned: vins_estimator/CMakeLists.txt:19).  The arithmetic below follows OpenCV's published algorithm:
  cornerMinEigenVal  = Sobel 3x3 (kernel scaled by 1/(255*4*3), FP32, BORDER_REFLECT_101) -> dx^2, dx*dy, dy^2 -> 3x3 un-normalised
                       box sum accumulated in FP64 and narrowed to FP32 -> (a+c) - sqrt((a-c)^2 + b^2) with a, c halved;
  goodFeaturesToTrack = masked maximum, THRESH_TOZERO at max*quality, 3x3 dilation, strict local maxima in rows/cols 1..n-2,
                       descending sort (ties: larger address first), greedy minimum-distance selection on a cell grid;
  cv::circle (filled) = midpoint circle, one horizontal run per row.
Pinning: the restatement is checked bit-for-bit against the cv2 4.13 wheel of this image with its SIMD paths switched off
(cv2.setUseOptimized(False): the scalar code is the specification; the SIMD build contracts some multiply-adds into FMAs and differs
from its own scalar path in ~15 % of the eigenvalue pixels by one ulp) -- tests/test_feature_oracle.py.  The reference itself holds
no fixture for these functions, so parity with the *reference's* OpenCV 3.x build stays unpinned.
"""
import numpy as np

f32 = np.float32


def cv_round(x):
    """saturate_cast<int>(float): round half to even (cvRound / lrint), what Point2f -> Point does"""
    return np.rint(np.asarray(x, np.float64)).astype(np.int64)


def circle_half_widths(r):
    """hw[|dy|] = half width of the filled run in the row |dy| away from the centre (imgproc/drawing.cpp Circle, fill branch)"""
    hw = np.full(r + 1, -1, np.int64)
    err, dx, dy, plus, minus = 0, r, 0, 1, (r << 1) - 1
    while dx >= dy:
        hw[dx] = max(hw[dx], dy)
        hw[dy] = max(hw[dy], dx)
        dy += 1
        err += plus
        plus += 2
        m = (1 if err <= 0 else 0) - 1
        err -= minus & m
        dx += m
        minus -= m & 2
    return hw


def paint_circle(mask, cx, cy, r, hw=None):
    h, w = mask.shape
    hw = circle_half_widths(r) if hw is None else hw
    for dy in range(-r, r + 1):
        y = cy + dy
        if 0 <= y < h:
            x0, x1 = max(0, cx - hw[abs(dy)]), min(w - 1, cx + hw[abs(dy)])
            if x0 <= x1:
                mask[y, x0:x1 + 1] = 0


def set_mask(width, height, pts, track_cnt, min_dist, base_mask=None):
    """feature_tracker.cpp:59-89.  Returns (mask, keep) where keep lists the surviving indices in their new order.
    Equal track counts keep their original relative order (std::sort leaves that order unspecified)."""
    mask = np.full((height, width), 255, np.uint8) if base_mask is None else base_mask.copy()
    pts = np.asarray(pts, f32).reshape(-1, 2)
    order = np.argsort(-np.asarray(track_cnt, np.int64), kind="stable")
    hw = circle_half_widths(int(min_dist))
    keep = []
    for i in order:
        x, y = int(cv_round(pts[i, 0])), int(cv_round(pts[i, 1]))
        if mask[y, x] == 255:
            keep.append(int(i))
            paint_circle(mask, x, y, int(min_dist), hw)
    return mask, np.array(keep, np.int32)


def corner_min_eigen_val(img):
    """cv::cornerMinEigenVal(img, eig, blockSize 3, ksize 3) on an 8-bit image, scalar code path"""
    h, w = img.shape
    scale = 1.0 / (4 * 3) / 255.0
    k1, k2 = f32(1 * scale), f32(2 * scale)
    p = np.pad(img.astype(f32), 1, mode="reflect")
    rdx = p[:, 2:] - p[:, :-2]                                      # row pass of d/dx: [-1 0 1]
    dx = k2 * rdx[1:-1] + k1 * (rdx[:-2] + rdx[2:])                 # symmetric column pass with the scaled [1 2 1]
    rdy = (k1 * p[:, :-2] + k2 * p[:, 1:-1]) + k1 * p[:, 2:]        # row pass of d/dy: scaled [1 2 1], accumulated left to right
    dy = rdy[2:] - rdy[:-2]                                         # column pass [-1 0 1]

    def box(c):                                                     # boxFilter(normalize=false): row sums, then column sums, in FP64
        q = np.pad(c.astype(np.float64), 1, mode="reflect")
        r = (q[:, :-2] + q[:, 1:-1]) + q[:, 2:]
        return ((r[:-2] + r[1:-1]) + r[2:]).astype(f32)

    a, b, c = box(dx * dx) * f32(0.5), box(dx * dy), box(dy * dy) * f32(0.5)
    return (a + c) - np.sqrt((a - c) * (a - c) + b * b)


def good_features_to_track(img, max_corners, quality, min_dist, mask=None, return_candidates=False):
    h, w = img.shape
    eig = corner_min_eigen_val(img)
    sel = eig[mask != 0] if mask is not None else eig.reshape(-1)
    if sel.size == 0:
        return np.zeros((0, 2), f32)
    thr = f32(np.float64(sel.max()) * quality)
    e = np.where(eig > thr, eig, f32(0))
    q = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = np.max([q[i:i + h, j:j + w] for i in range(3) for j in range(3)], axis=0)
    ok = (e != 0) & (e == dil)
    if mask is not None:
        ok &= mask != 0
    ok[0, :] = ok[-1, :] = False
    ok[:, 0] = ok[:, -1] = False
    ys, xs = np.nonzero(ok)
    vals, ofs = e[ys, xs], ys * w + xs
    order = np.lexsort((-ofs, -vals.astype(np.float64)))
    out = []
    if min_dist >= 1:
        cell = int(cv_round(min_dist))
        gw, gh = (w + cell - 1) // cell, (h + cell - 1) // cell
        grid = [[] for _ in range(gw * gh)]
        md2 = float(min_dist) ** 2
        for i in order:
            x, y = int(xs[i]), int(ys[i])
            xc, yc = x // cell, y // cell
            good = True
            for yy in range(max(0, yc - 1), min(gh - 1, yc + 1) + 1):
                for xx in range(max(0, xc - 1), min(gw - 1, xc + 1) + 1):
                    for (px, py) in grid[yy * gw + xx]:
                        if (x - px) ** 2 + (y - py) ** 2 < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid[yc * gw + xc].append((x, y))
                out.append((x, y))
                if max_corners > 0 and len(out) == max_corners:
                    break
    else:
        for i in order:
            out.append((int(xs[i]), int(ys[i])))
            if max_corners > 0 and len(out) == max_corners:
                break
    res = np.array(out, f32).reshape(-1, 2)
    return (res, len(ys)) if return_candidates else res


def detect(img, pts, track_cnt, max_cnt, min_dist, quality=0.01, base_mask=None):
    """setMask() followed by goodFeaturesToTrack(), exactly as trackImage() chains them (feature_tracker.cpp:175-200)"""
    h, w = img.shape
    mask, keep = set_mask(w, h, pts, track_cnt, min_dist, base_mask)
    n_new = max_cnt - len(keep)
    new = good_features_to_track(img, n_new, quality, float(min_dist), mask) if n_new > 0 else np.zeros((0, 2), f32)
    return mask, keep, new

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


def set_mask(width, height, pts, track_cnt, min_dist, base_mask=None, order_fn=None):
    """feature_tracker.cpp:59-89.  Returns (mask, keep) where keep lists the surviving indices in their new order.
    Equal track counts keep their original relative order (std::sort leaves that order unspecified); order_fn(track_cnt) -> visiting order
    substitutes another valid order (the pinning test passes the one this libstdc++'s std::sort produces)."""
    mask = np.full((height, width), 255, np.uint8) if base_mask is None else base_mask.copy()
    pts = np.asarray(pts, f32).reshape(-1, 2)
    order = np.argsort(-np.asarray(track_cnt, np.int64), kind="stable") if order_fn is None else order_fn(track_cnt)
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


# ------------------------------------------------------------------------------------------------ FeatureTracker::trackImage()
def lift_projective(cam, pts):
    """PinholeCamera::liftProjective with the recursive distortion model, n = 8 (camera_models/src/camera_models/PinholeCamera.cc:450-517),
    narrowed to cv::Point2f as FeatureTracker::undistortedPts does (feature_tracker.cpp:606-617).  cam = (fx, fy, cx, cy, k1, k2, p1, p2)."""
    fx, fy, cx, cy, k1, k2, p1, p2 = (float(v) for v in cam)
    p = np.asarray(pts, f32).reshape(-1, 2)
    mxd, myd = (1.0 / fx) * p[:, 0].astype(np.float64) + (-cx / fx), (1.0 / fy) * p[:, 1].astype(np.float64) + (-cy / fy)
    mx, my = mxd.copy(), myd.copy()
    if not (k1 == 0.0 and k2 == 0.0 and p1 == 0.0 and p2 == 0.0):
        for it in range(8):
            x, y = (mxd, myd) if it == 0 else (mx, my)
            mx2, my2, mxy = x * x, y * y, x * y
            rho2 = mx2 + my2
            rad = k1 * rho2 + k2 * rho2 * rho2
            dx = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2)
            dy = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2)
            mx, my = mxd - dx, myd - dy
    return np.column_stack([mx, my]).astype(f32)


class FeatureTrackerRef:
    """FeatureTracker (featureTracker/feature_tracker.h / .cpp) restated line by line, one session.  The two OpenCV calls stay OpenCV
    calls -- cv2.calcOpticalFlowPyrLK is the very routine the reference links -- so this class is the reference's trackImage() with
    its own third-party library underneath; setMask / goodFeaturesToTrack go through the restatement above (bit-identical to cv2's
    scalar path, tests/test_feature_oracle.py) or through cv2 itself (use_cv_detector=True)."""

    def __init__(self, cam0, cam1=None, max_cnt=150, min_dist=30, flow_back=True, use_cv_detector=False, order_fn=None):
        import cv2
        self.cv2 = cv2
        self.order_fn = order_fn
        self.cam, self.stereo = (cam0, cam1), cam1 is not None
        self.MAX_CNT, self.MIN_DIST, self.FLOW_BACK, self.use_cv = max_cnt, min_dist, flow_back, use_cv_detector
        self.n_id = 0
        self.prev_img = None
        self.prev_pts = np.zeros((0, 2), f32)
        self.ids, self.track_cnt = [], []
        self.prev_un_pts_map, self.prev_un_right_pts_map = {}, {}
        self.prev_time = 0.0
        self.hasPrediction, self.predict_pts = False, None
        self.stats = {"predicted": 0, "repeated": 0}          # how often the hasPrediction branch / its full-pyramid repeat ran

    def set_prediction(self, predict_pts):
        """FeatureTracker::setPrediction leaves predict_pts aligned with prev_pts (feature_tracker.cpp:715-736)"""
        self.hasPrediction, self.predict_pts = True, np.asarray(predict_pts, f32).reshape(-1, 2).copy()

    def _in_border(self, p):                                              # feature_tracker.cpp:19-25
        x, y = int(cv_round(p[0])), int(cv_round(p[1]))
        return 1 <= x < self.col - 1 and 1 <= y < self.row - 1

    def _lk(self, a, b, pts, init=None, max_level=3):
        cv2 = self.cv2
        p = np.ascontiguousarray(pts, f32).reshape(-1, 1, 2)
        if init is None:
            q, st, _ = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=(21, 21), maxLevel=max_level)
        else:
            q, st, _ = cv2.calcOpticalFlowPyrLK(a, b, p, np.ascontiguousarray(init, f32).reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=max_level,
                                                criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01), flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        return q.reshape(-1, 2), st.reshape(-1).astype(bool)

    def _velocity(self, ids, un, prev_map, dt):                           # ptsVelocity, :619-657
        vel = np.zeros((len(ids), 2), f32)
        if prev_map:
            for i, fid in enumerate(ids):
                if fid in prev_map:
                    vel[i] = ((un[i] - prev_map[fid]).astype(np.float64) / dt).astype(f32)      # float difference, double division
        return vel

    def track_image(self, cur_time, img, img1=None):
        """-> (ids, track_cnt, cur_pts, cur_un_pts, pts_velocity, ids_right, cur_right_pts, cur_un_right_pts, right_pts_velocity)"""
        self.row, self.col = img.shape
        cur_pts = np.zeros((0, 2), f32)
        if len(self.prev_pts) > 0:                                        # :117-172
            if self.hasPrediction:
                cur_pts, status = self._lk(self.prev_img, img, self.prev_pts, self.predict_pts, 1)
                self.stats["predicted"] += 1
                if status.sum() < 10:
                    self.stats["repeated"] += 1
                    cur_pts, status = self._lk(self.prev_img, img, self.prev_pts)
            else:
                cur_pts, status = self._lk(self.prev_img, img, self.prev_pts)
            if self.FLOW_BACK:
                rev, rst = self._lk(img, self.prev_img, cur_pts, self.prev_pts, 1)
                d = (self.prev_pts - rev).astype(np.float64)
                status = status & rst & (np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2) <= 0.5)
            status = status & np.array([self._in_border(p) for p in cur_pts], bool)
            cur_pts = cur_pts[status]
            self.ids = [i for i, s in zip(self.ids, status) if s]
            self.track_cnt = [c for c, s in zip(self.track_cnt, status) if s]
        self.track_cnt = [c + 1 for c in self.track_cnt]                  # :174-175
        # setMask (:59-89) rebuilds cur_pts / ids / track_cnt in visiting order; goodFeaturesToTrack tops up (:181-204)
        cnt = np.asarray(self.track_cnt, np.int64)
        mask, keep = set_mask(self.col, self.row, cur_pts, cnt, self.MIN_DIST, order_fn=self.order_fn)
        cur_pts = cur_pts[keep] if len(keep) else np.zeros((0, 2), f32)
        self.ids = [self.ids[k] for k in keep]
        self.track_cnt = [self.track_cnt[k] for k in keep]
        n_max_cnt = self.MAX_CNT - len(cur_pts)
        if n_max_cnt > 0:
            if self.use_cv:
                n_pts = self.cv2.goodFeaturesToTrack(img, n_max_cnt, 0.01, self.MIN_DIST, mask=mask)
                n_pts = np.zeros((0, 2), f32) if n_pts is None else n_pts.reshape(-1, 2)
            else:
                n_pts = good_features_to_track(img, n_max_cnt, 0.01, float(self.MIN_DIST), mask)
        else:
            n_pts = np.zeros((0, 2), f32)
        for p in n_pts:
            self.ids.append(self.n_id); self.n_id += 1
            self.track_cnt.append(1)
        cur_pts = np.concatenate([cur_pts, n_pts.astype(f32)]).astype(f32)
        dt = cur_time - self.prev_time
        cur_un_pts = lift_projective(self.cam[0], cur_pts)                # :207-208
        pts_velocity = self._velocity(self.ids, cur_un_pts, self.prev_un_pts_map, dt)
        cur_un_pts_map = {fid: cur_un_pts[i] for i, fid in enumerate(self.ids)}
        ids_right, cur_right_pts = [], np.zeros((0, 2), f32)
        cur_un_right_pts, right_pts_velocity = np.zeros((0, 2), f32), np.zeros((0, 2), f32)
        if img1 is not None and self.stereo:                              # :210-271
            cur_un_right_pts_map = {}
            if len(cur_pts) > 0:
                right, status = self._lk(img, img1, cur_pts)
                if self.FLOW_BACK:
                    rev, rst = self._lk(img1, img, right)
                    d = (cur_pts - rev).astype(np.float64)
                    status = status & rst & np.array([self._in_border(p) for p in right], bool) & (np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2) <= 0.5)
                cur_right_pts = right[status]
                ids_right = [i for i, s in zip(self.ids, status) if s]
                cur_un_right_pts = lift_projective(self.cam[1], cur_right_pts)
                right_pts_velocity = self._velocity(ids_right, cur_un_right_pts, self.prev_un_right_pts_map, dt)
                cur_un_right_pts_map = {fid: cur_un_right_pts[i] for i, fid in enumerate(ids_right)}
            self.prev_un_right_pts_map = cur_un_right_pts_map
        self.prev_img, self.prev_pts, self.prev_un_pts_map, self.prev_time = img, cur_pts, cur_un_pts_map, cur_time      # :296-301
        self.hasPrediction = False
        return (np.asarray(self.ids, np.int32), np.asarray(self.track_cnt, np.int32), cur_pts, cur_un_pts, pts_velocity,
                np.asarray(ids_right, np.int32), cur_right_pts, cur_un_right_pts, right_pts_velocity)

"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libviw_ref.so: the reference's OWN factor / manifold / pre-integration
sources (vins_estimator/src/factor/*.cpp, *.h), compiled unmodified from /root/reference against the header stand-ins of
oracle/refshim/ (`make -C oracle ref`).  Used by tests/test_reference_factors.py to pin the restated oracle (and through it the CUDA
library) to the reference code itself.  The library is git-ignored and rebuilt where /root/reference exists."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libviw_ref.so")
REF_SRC = "/root/reference/vins_estimator/src"
_lib = None
c_double_p = C.POINTER(C.c_double)


def available():
    return os.path.exists(LIB) or os.path.isdir(REF_SRC)


def lib():
    global _lib
    if _lib is None:
        if os.path.isdir(REF_SRC):
            subprocess.check_call(["make", "-C", HERE, "-s", "ref"])
        _lib = C.CDLL(LIB)
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def factor_evaluate(ftype, globals_, consts, params, sizes, nres, want_jac=True, null_jac=()):
    P = [np.ascontiguousarray(p, np.float64) for p in params]
    pp = (c_double_p * len(P))(*[_dp(p) for p in P])
    res = np.zeros(nres)
    cst = np.ascontiguousarray(consts, np.float64) if consts is not None else None
    jacs = [None if (not want_jac or i in null_jac) else np.zeros((nres, s)) for i, s in enumerate(sizes)]
    jp = (c_double_p * len(P))(*[(_dp(j) if j is not None else None) for j in jacs]) if want_jac else None
    rc = lib().ref_factor_evaluate(C.c_int(ftype), C.byref(globals_), _dp(cst) if cst is not None else None, pp, _dp(res), jp)
    if rc:
        raise RuntimeError("ref_factor_evaluate rc=%d" % rc)
    return res, jacs


def imu_preintegrate(dt, acc, gyr, ba, bg, noise):
    dt, acc, gyr = (np.ascontiguousarray(a, np.float64) for a in (dt, acc, gyr))
    ba, bg, noise = (np.ascontiguousarray(a, np.float64) for a in (ba, bg, noise))
    rec = np.zeros(287)
    lib().ref_imu_preintegrate(C.c_int(len(dt)), _dp(dt), _dp(acc), _dp(gyr), _dp(ba), _dp(bg), _dp(noise), _dp(rec))
    return rec


def wheel_preintegrate(dt, vel, gyr, s, td, noise):
    dt, vel, gyr = (np.ascontiguousarray(a, np.float64) for a in (dt, vel, gyr))
    s, noise = np.ascontiguousarray(s, np.float64), np.ascontiguousarray(noise, np.float64)
    rec = np.zeros(78)
    lib().ref_wheel_preintegrate(C.c_int(len(dt)), _dp(dt), _dp(vel), _dp(gyr), _dp(s), C.c_double(td), _dp(noise), _dp(rec))
    return rec


def manifold(kind, mask, x, delta, want_jac=True):
    """kind 0 PoseLocalParameterization, 1 PoseSubsetParameterization, 2 OrientationSubsetParameterization"""
    gs, ls = (7, 6) if kind < 2 else (4, 3)
    x, delta = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(delta, np.float64)
    out, jac = np.zeros(gs), np.zeros((gs, ls))
    rc = lib().ref_manifold(C.c_int(kind), C.c_uint(mask), _dp(x), _dp(delta), _dp(out), _dp(jac) if want_jac else None)
    if rc:
        raise RuntimeError("ref_manifold rc=%d" % rc)
    return out, jac


def prior_evaluate(prior, state, want_jac=True):
    """MarginalizationFactor::Evaluate of the reference on a viwb_prior (abi.PriorData)"""
    st = np.ascontiguousarray(state, np.float64)
    res = np.zeros(prior.n)
    jac = np.zeros((prior.n, 207)) if want_jac else None
    rc = lib().ref_prior_evaluate(C.byref(prior.c), _dp(st), _dp(res), _dp(jac) if want_jac else None)
    if rc:
        raise RuntimeError("ref_prior_evaluate rc=%d" % rc)
    return res, jac


def marginalize(problem, state, flag):
    """MarginalizationInfo::preMarginalize + marginalize of the reference on the factors Estimator::optimization() hands over.
    Returns dict(m, n, blocks=[(block id BEFORE the addr_shift remap, column offset)], J, r)."""
    st = np.ascontiguousarray(state, np.float64)
    cap = 256
    mn = (C.c_int32 * 3)()
    bid, bidx = (C.c_int32 * 32)(), (C.c_int32 * 32)()
    J, r = np.zeros(cap * cap), np.zeros(cap)
    rc = lib().ref_marginalize(C.byref(problem.c), _dp(st), C.c_int(flag), mn, bid, bidx, _dp(J), _dp(r))
    if rc:
        raise RuntimeError("ref_marginalize rc=%d" % rc)
    n = mn[1]
    return {"m": mn[0], "n": n, "blocks": [(bid[k], bidx[k]) for k in range(mn[2])], "J": J[: n * n].reshape(n, n).copy(), "r": r[:n].copy()}


def triangulate(state, stereo, frame, pt0, pt1, init_depth=5.0):
    """FeatureManager::triangulate (feature_manager.cpp:309-438) of the reference on two-observation features"""
    st = np.ascontiguousarray(state, np.float64)
    s_, f_ = np.ascontiguousarray(stereo, np.int32), np.ascontiguousarray(frame, np.int32)
    a, b = np.ascontiguousarray(pt0, np.float64), np.ascontiguousarray(pt1, np.float64)
    out = np.zeros(len(s_))
    ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib().ref_triangulate(_dp(st), C.c_int(len(s_)), ip(s_), ip(f_), _dp(a), _dp(b), C.c_double(init_depth), _dp(out))
    assert rc == 0
    return out


def shift_depth(uv, depth, marg_R, marg_P, new_R, new_P, init_depth=5.0):
    """FeatureManager::removeBackShiftDepth (feature_manager.cpp:457-495) of the reference"""
    arrs = [np.ascontiguousarray(x, np.float64) for x in (uv, depth, marg_R, marg_P, new_R, new_P)]
    out = np.zeros(len(arrs[1]))
    rc = lib().ref_shift_depth(C.c_int(len(out)), *[_dp(x) for x in arrs], C.c_double(init_depth), _dp(out))
    assert rc == 0
    return out


def undistorted_pts(cam, width, height, pts):
    """FeatureTracker::undistortedPts over the reference's PinholeCamera::liftProjective (camera_models/src/camera_models/PinholeCamera.cc)"""
    c = np.ascontiguousarray(cam, np.float64)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(p)
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    lib().ref_undistorted_pts(_dp(c), C.c_int(width), C.c_int(height), C.c_int(len(p)), fp(p), fp(out))
    return out


def space_to_plane(cam, width, height, P):
    c = np.ascontiguousarray(cam, np.float64)
    P = np.ascontiguousarray(P, np.float64).reshape(-1, 3)
    uv = np.zeros((len(P), 2))
    lib().ref_space_to_plane(_dp(c), C.c_int(width), C.c_int(height), C.c_int(len(P)), _dp(P), _dp(uv))
    return uv


def estimator_optimization(problem, state, state_solved, flag):
    """The reference's own Estimator::optimization() (estimator/estimator.cpp compiled unmodified) on the window `state`, with ceres::Solve
    replaced by a hook that checks the assembled problem against the tables and plays `state_solved` back.  Returns a dict: state_out (the window
    after double2vector + vector2double), the new prior (m, n, blocks after the address shift, J, r) and the assembly record."""
    st = np.ascontiguousarray(state, np.float64)
    so = np.ascontiguousarray(state_solved, np.float64)
    out = np.zeros_like(st)
    cap = 256
    mn, bid, bidx, rec = (C.c_int32 * 3)(), (C.c_int32 * 32)(), (C.c_int32 * 32)(), (C.c_int32 * 11)()
    J, r = np.zeros(cap * cap), np.zeros(cap)
    rc = lib().ref_estimator_optimization(C.byref(problem.c), _dp(st), _dp(so), C.c_int(flag), _dp(out), mn, bid, bidx, _dp(J), _dp(r), rec)
    if rc:
        raise RuntimeError("ref_estimator_optimization rc=%d" % rc)
    n = mn[1]
    names = ("prior", "imu", "wheel", "plane", "proj_2f1c", "proj_2f2c", "proj_1f2c", "parameter_blocks", "structure_mismatches", "vector2double_mismatches", "visual_row_mismatches")
    return {"state": out, "m": mn[0], "n": n, "blocks": [(bid[k], bidx[k]) for k in range(mn[2])], "J": J[: n * n].reshape(n, n).copy(), "r": r[:n].copy(),
            "record": dict(zip(names, [int(v) for v in rec]))}


def estimator_outliers(problem, state):
    """Estimator::outliersRejection (estimator.cpp:2127-2185) of the reference"""
    st = np.ascontiguousarray(state, np.float64)
    out = np.zeros(max(problem.num_landmarks, 1), np.uint8)
    rc = lib().ref_estimator_outliers(C.byref(problem.c), _dp(st), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc:
        raise RuntimeError("ref_estimator_outliers rc=%d" % rc)
    return out[: problem.num_landmarks]


def estimator_optimization_with(problem, state, solve, flag):
    """As estimator_optimization, but ceres::Solve hands the window to `solve(state) -> solved state` (e.g. the library under test): the
    reference's unmodified Estimator::optimization() then runs on that backend."""
    from viwb import abi
    st = np.ascontiguousarray(state, np.float64)
    out = np.zeros_like(st)
    size = len(st)
    CB = C.CFUNCTYPE(C.c_int, C.POINTER(abi.Problem), C.POINTER(C.c_double))
    err = []

    def cb(_pp, sp):
        try:
            window = np.ctypeslib.as_array(sp, shape=(size,))
            window[:] = solve(window.copy())
            return 0
        except Exception as ex:            # noqa: BLE001 -- reported to the caller below
            err.append(ex)
            return 1
    cap = 256
    mn, bid, bidx, rec = (C.c_int32 * 3)(), (C.c_int32 * 32)(), (C.c_int32 * 32)(), (C.c_int32 * 11)()
    J, r = np.zeros(cap * cap), np.zeros(cap)
    rc = lib().ref_estimator_optimization_with(C.byref(problem.c), _dp(st), CB(cb), C.c_int(flag), _dp(out), mn, bid, bidx, _dp(J), _dp(r), rec)
    if err:
        raise err[0]
    if rc:
        raise RuntimeError("ref_estimator_optimization_with rc=%d" % rc)
    n = mn[1]
    names = ("prior", "imu", "wheel", "plane", "proj_2f1c", "proj_2f2c", "proj_1f2c", "parameter_blocks", "structure_mismatches", "vector2double_mismatches", "visual_row_mismatches")
    return {"state": out, "m": mn[0], "n": n, "blocks": [(bid[k], bidx[k]) for k in range(mn[2])], "J": J[: n * n].reshape(n, n).copy(), "r": r[:n].copy(),
            "record": dict(zip(names, [int(v) for v in rec]))}


def visual_imu_alignment(R, T, dts, accs, gyrs, noise, bg0, wheel_rec, tic, rio, tio, gvec):
    """VisualIMUAlignment of the reference (initial/initial_aligment.cpp compiled unmodified): solveGyroscopeBias with its repropagation, then
    LinearAlignment[WithWheel] + RefineGravity[WithWheel].  Returns dict(ok, delta_bg, imu (repropagated records), g, x)."""
    F = len(R)
    R = np.ascontiguousarray(R, np.float64); T = np.ascontiguousarray(T, np.float64)
    counts = np.array([len(d) for d in dts], np.int32)
    dt = np.ascontiguousarray(np.concatenate(dts), np.float64)
    acc = np.ascontiguousarray(np.concatenate(accs), np.float64); gyr = np.ascontiguousarray(np.concatenate(gyrs), np.float64)
    noise, bg0, tic, gvec = (np.ascontiguousarray(a, np.float64) for a in (noise, bg0, tic, gvec))
    w = np.ascontiguousarray(wheel_rec, np.float64) if wheel_rec is not None else None
    rio = np.ascontiguousarray(rio, np.float64) if rio is not None else None
    tio = np.ascontiguousarray(tio, np.float64) if tio is not None else None
    dbg, rec, g, x, xs = np.zeros(3), np.zeros((F - 1, 287)), np.zeros(3), np.zeros(3 * F + 4), C.c_int32(0)
    ok = lib().ref_visual_imu_alignment(C.c_int(F), _dp(R), _dp(T), counts.ctypes.data_as(C.POINTER(C.c_int32)), _dp(dt), _dp(acc), _dp(gyr), _dp(noise), _dp(bg0),
                                        _dp(w) if w is not None else None, _dp(tic), _dp(rio) if rio is not None else None, _dp(tio) if tio is not None else None, _dp(gvec),
                                        _dp(dbg), _dp(rec), _dp(g), _dp(x), C.byref(xs))
    return {"ok": bool(ok), "delta_bg": dbg, "imu": rec, "g": g, "x": x[: xs.value].copy()}


# ---- third build: the reference's estimator.cpp compiled against the PRODUCT's ceres shim (viw-fusion_b200/host + viwb_reference_adapter.h)
PRODUCT_LIB = os.path.join(HERE, "_ref", "libviw_ref_product.so")
_plib = {}


PRODUCT_DEV_LIB = os.path.join(HERE, "_ref", "libviw_ref_product_dev.so")      # the same with the marginalization on the device (marginalization_factor_device.cpp)
_under_test = []


def product_lib(under_test_path, dev=False):
    """`under_test_path`: the libviwb*.so the shim's viwb_* calls are to land in (loaded RTLD_GLOBAL first: the checker library leaves them undefined)."""
    key = (under_test_path, dev)
    if key not in _plib:
        if _under_test and _under_test[0] != under_test_path:
            raise RuntimeError("one library under test per process")
        if os.path.isdir(REF_SRC):
            subprocess.check_call(["make", "-C", HERE, "-s", "ref_product_dev" if dev else "ref_product"])
        if not _under_test:
            C.CDLL(under_test_path, mode=C.RTLD_GLOBAL)
            _under_test.append(under_test_path)
        _plib[key] = C.CDLL(PRODUCT_DEV_LIB if dev else PRODUCT_LIB)
    return _plib[key]


def product_available(dev=False):
    return os.path.exists(PRODUCT_DEV_LIB if dev else PRODUCT_LIB) or os.path.isdir(REF_SRC)


def estimator_optimization_on_product_shim(under_test_path, problem, state, flag, dev=False):
    """Estimator::optimization() of the reference, compiled unmodified, with <ceres/ceres.h> = the product shim: ceres::Problem / ceres::Solve are
    the product's host code, the reference's own factor objects are lowered by the product's adapter, the solve runs in the library under test;
    the marginalization that follows is the reference's own CPU code on its own factor classes."""
    st = np.ascontiguousarray(state, np.float64)
    out = np.zeros_like(st)
    cap = 256
    mn, bid, bidx, rec = (C.c_int32 * 3)(), (C.c_int32 * 32)(), (C.c_int32 * 32)(), (C.c_int32 * 11)()
    J, r = np.zeros(cap * cap), np.zeros(cap)
    rc = product_lib(under_test_path, dev).ref_estimator_optimization(C.byref(problem.c), _dp(st), None, C.c_int(flag), _dp(out), mn, bid, bidx, _dp(J), _dp(r), rec)
    if rc:
        raise RuntimeError("ref_estimator_optimization (product shim) rc=%d" % rc)
    n = mn[1]
    return {"state": out, "m": mn[0], "n": n, "blocks": [(bid[k], bidx[k]) for k in range(mn[2])], "J": J[: n * n].reshape(n, n).copy(), "r": r[:n].copy()}


def prior_factor_digest_on_product_shim(under_test_path, problem, state, flag, dev=False):
    """As estimator_optimization_on_product_shim, followed by MarginalizationFactor::Evaluate of the NEW prior at a perturbed copy of its blocks:
    (|res|^2, sum_b |J_b^T res|^2, n) -- the reference's own Evaluate (dev=False) or the device translation unit's (dev=True)."""
    st = np.ascontiguousarray(state, np.float64)
    out = np.zeros_like(st)
    cap = 256
    mn, bid, bidx, rec = (C.c_int32 * 3)(), (C.c_int32 * 32)(), (C.c_int32 * 32)(), (C.c_int32 * 11)()
    J, r, dig = np.zeros(cap * cap), np.zeros(cap), np.zeros(3)
    rc = product_lib(under_test_path, dev).ref_estimator_optimization_prior_eval(C.byref(problem.c), _dp(st), C.c_int(flag), _dp(out), mn, bid, bidx, _dp(J), _dp(r), rec, _dp(dig))
    if rc:
        raise RuntimeError("ref_estimator_optimization_prior_eval rc=%d" % rc)
    return dig


# ---- the reference's FeatureTracker (featureTracker/feature_tracker.cpp, compiled unmodified) with its OpenCV calls answered by the real cv2
_LK = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_ubyte),
                  C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int)
_GFTT = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.c_int, C.c_double, C.c_double, C.POINTER(C.c_float), C.c_int)
_CIRCLE = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_cv_keep = []
_cv_backend = [None]          # None: the real cv2 answers FeatureTracker's OpenCV calls; a viwb Context: the library under test does


def use_backend(ctx):
    """ctx = a viwb.lib.Context: cv::calcOpticalFlowPyrLK and cv::goodFeaturesToTrack inside the reference's tracker code are answered by
    viwb_lk_track / viwb_good_features_to_track (literally); None: by cv2.  cv::circle stays cv2's drawing."""
    _cv_backend[0] = ctx


def _install_cv_callbacks():
    import cv2
    if _cv_keep:
        return

    def img(ptr, rows, cols):
        return np.ctypeslib.as_array(ptr, shape=(rows, cols))

    def lk(prev, nxt, rows, cols, n, p0, p1, status, err, win, max_level, crit_count, crit_eps, flags):
        a, b = img(prev, rows, cols), img(nxt, rows, cols)
        pts0 = np.ctypeslib.as_array(p0, shape=(n, 2)).reshape(-1, 1, 2).copy()
        pts1 = np.ctypeslib.as_array(p1, shape=(n, 2))
        if _cv_backend[0] is not None:
            q, st, er = _cv_backend[0].lk_track(a, b, pts0.reshape(-1, 2), pts1.copy() if flags & 4 else None, max_level=max_level, max_iter=crit_count, eps=crit_eps, flags=flags)
            pts1[:] = q
            np.ctypeslib.as_array(status, shape=(n,))[:] = st
            np.ctypeslib.as_array(err, shape=(n,))[:] = er
            return
        q, st, er = cv2.calcOpticalFlowPyrLK(a, b, pts0, pts1.reshape(-1, 1, 2).copy() if flags & 4 else None, winSize=(win, win), maxLevel=max_level,
                                             criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, crit_count, crit_eps), flags=flags)
        pts1[:] = q.reshape(-1, 2)
        np.ctypeslib.as_array(status, shape=(n,))[:] = st.reshape(-1)
        np.ctypeslib.as_array(err, shape=(n,))[:] = er.reshape(-1)

    def gftt(im, rows, cols, mask, max_corners, quality, min_dist, out, cap):
        m = img(mask, rows, cols) if mask else None
        if _cv_backend[0] is not None:
            c = _cv_backend[0].good_features_to_track(img(im, rows, cols), max_corners, quality, min_dist, mask=m)
        else:
            c = cv2.goodFeaturesToTrack(img(im, rows, cols), max_corners, quality, min_dist, mask=m)
        if c is None:
            return 0
        c = c.reshape(-1, 2)[:cap]
        np.ctypeslib.as_array(out, shape=(cap, 2))[: len(c)] = c
        return len(c)

    def circle(im, rows, cols, cx, cy, radius, color, thickness):
        cv2.circle(img(im, rows, cols), (cx, cy), radius, color, thickness)
    cbs = (_LK(lk), _GFTT(gftt), _CIRCLE(circle))
    _cv_keep.extend(cbs)
    lib().ref_set_cv_callbacks(*cbs)


class ReferenceFeatureTracker:
    """FeatureTracker of the reference (one session): track_image(t, left, right) -> (ids, track_cnt, feat[n,6], ids_right, feat_right[m,6])"""

    def __init__(self, cam0, cam1, width, height, max_cnt, min_dist, flow_back=True):
        _install_cv_callbacks()
        L = lib()
        L.ref_tracker_create.restype = C.c_void_p
        c0 = np.ascontiguousarray(cam0, np.float64)
        c1 = None if cam1 is None else np.ascontiguousarray(cam1, np.float64)
        self.cap = max(2 * max_cnt, 64)
        self.h = C.c_void_p(L.ref_tracker_create(_dp(c0), _dp(c1) if c1 is not None else None, C.c_int(width), C.c_int(height), C.c_int(max_cnt), C.c_int(min_dist),
                                                 C.c_int(1 if flow_back else 0)))

    def set_prediction(self, pts):
        p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        rc = lib().ref_tracker_set_prediction(self.h, C.c_int(len(p)), p.ctypes.data_as(C.POINTER(C.c_float)))
        assert rc == 0, "prediction must have one point per previous point"

    def track_image(self, t, left, right=None):
        cap = self.cap
        nl, nr = C.c_int32(), C.c_int32()
        ids, cnt, ids_r = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        feat, feat_r = np.zeros((cap, 6), np.float32), np.zeros((cap, 6), np.float32)
        a = np.ascontiguousarray(left, np.uint8)
        b = None if right is None else np.ascontiguousarray(right, np.uint8)
        up = lambda x: x.ctypes.data_as(C.POINTER(C.c_ubyte))
        ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
        fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        rc = lib().ref_tracker_track(self.h, C.c_double(t), up(a), up(b) if b is not None else None, C.c_int(cap), C.byref(nl), ip(ids), ip(cnt), fp(feat),
                                     C.byref(nr), ip(ids_r), fp(feat_r))
        if rc:
            raise RuntimeError("ref_tracker_track rc=%d" % rc)
        return ids[: nl.value].copy(), cnt[: nl.value].copy(), feat[: nl.value].copy(), ids_r[: nr.value].copy(), feat_r[: nr.value].copy()


def std_sort_order(track_cnt):
    """visiting order of FeatureTracker::setMask under this libstdc++'s std::sort (equal counts: unspecified by the standard)"""
    c = np.ascontiguousarray(track_cnt, np.int32)
    out = np.zeros(len(c), np.int32)
    ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
    lib().ref_std_sort_order(C.c_int(len(c)), ip(c), ip(out))
    return out

"""ctypes wrapper of the CPU oracle (oracle/libviw_oracle.so).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.  Factor / manifold / pre-integration / marginalization math pinned to the reference's own compiled sources (oracle/_ref,
tests/test_reference_factors.py); the Ceres solve itself is PARITY UNPINNED (see oracle/viw_oracle.h).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "viw-fusion_b200", "python"))
from viwb import abi  # noqa: E402  (data-format definitions only)

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libviw_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("vo_factors.c", "vo_solver.c", "vo_marg.c", "viw_oracle.h", "vo_math.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "viwb.h"))
    stale = (not os.path.exists(so)) or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return so


class TraceEntry(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32), ("reused", C.c_int32),
                ("cost", C.c_double), ("cost_change", C.c_double), ("model_cost_change", C.c_double),
                ("relative_decrease", C.c_double), ("step_norm", C.c_double), ("x_norm", C.c_double),
                ("radius", C.c_double), ("mu", C.c_double), ("gradient_max_norm", C.c_double),
                ("dogleg_step_norm", C.c_double), ("alpha", C.c_double)]


class Trace(C.Structure):
    _fields_ = [("count", C.c_int32), ("e", TraceEntry * 64)]

    def rows(self):
        return [{k: getattr(self.e[i], k) for k, _ in TraceEntry._fields_} for i in range(self.count)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.vo_window_solve.restype = C.c_int
        _LIB.vo_cost.restype = C.c_int
    return _LIB


def _dp(a):
    return a.ctypes.data_as(abi.c_double_p)


def factor_evaluate(ftype, globals_, consts, params, want_jac=True, null_jac=()):
    """CostFunction::Evaluate.  Returns (residuals, [jacobian per block (rows x global size) or None])."""
    sizes = abi.FACTOR_BLOCK_SIZES[ftype]
    nres = abi.FACTOR_RESIDUALS[ftype]
    P = [np.ascontiguousarray(p, np.float64) for p in params]
    pp = (abi.c_double_p * len(P))(*[_dp(p) for p in P])
    res = np.zeros(nres)
    cst = np.ascontiguousarray(consts, np.float64) if consts is not None else None
    jacs = [None if (not want_jac or i in null_jac) else np.zeros((nres, s)) for i, s in enumerate(sizes)]
    jp = (abi.c_double_p * len(P))(*[(_dp(j) if j is not None else None) for j in jacs]) if want_jac else None
    rc = lib().vo_factor_evaluate(C.c_int(ftype), C.byref(globals_), _dp(cst) if cst is not None else None, pp, _dp(res), jp)
    if rc:
        raise RuntimeError("vo_factor_evaluate rc=%d" % rc)
    return res, jacs


def prior_evaluate(prior, state, want_jac=True):
    st = np.ascontiguousarray(state, np.float64)
    res = np.zeros(prior.n)
    jac = np.zeros((prior.n, abi.STATE_FIXED)) if want_jac else None
    lib().vo_prior_evaluate(C.byref(prior.c), _dp(st), _dp(res), _dp(jac) if want_jac else None)
    return res, jac


def cost(problem, state, want_residuals=False):
    st = np.ascontiguousarray(state, np.float64)
    c = C.c_double()
    nrows = 2 * len(problem.vis_type) + 15 * len(problem.imu_frame_i) + 6 * len(problem.wheel_frame_i) + 3 * len(problem.plane_frame)
    if problem.prior is not None and problem.prior.valid:
        nrows += problem.prior.n
    res = np.zeros(nrows) if want_residuals else None
    rc = lib().vo_cost(C.byref(problem.c), _dp(st), C.byref(c), _dp(res) if want_residuals else None)
    if rc < 0:
        raise RuntimeError("vo_cost rc=%d" % rc)
    return (c.value, res) if want_residuals else c.value


def normal_equations(problem, state):
    st = np.ascontiguousarray(state, np.float64)
    T = abi.TANGENT_FIXED
    H, g = np.zeros((T, T)), np.zeros(T)
    lm = np.zeros((max(problem.num_landmarks, 1), 82))
    c = C.c_double()
    rc = lib().vo_normal_equations(C.byref(problem.c), _dp(st), _dp(H), _dp(g), _dp(lm), C.byref(c))
    if rc:
        raise RuntimeError("vo_normal_equations rc=%d" % rc)
    return H, g, lm[: problem.num_landmarks], c.value


def window_solve(problem, state, options=None, want_trace=False):
    st = np.array(state, np.float64, copy=True)
    opt = options if options is not None else abi.default_options()
    summ = abi.Summary()
    tr = Trace()
    rc = lib().vo_window_solve(C.byref(problem.c), _dp(st), C.byref(opt), C.byref(summ), C.byref(tr))
    if rc:
        raise RuntimeError("vo_window_solve rc=%d" % rc)
    return (st, summ, tr.rows()) if want_trace else (st, summ)


def state_plus(problem, state, delta):
    st = np.ascontiguousarray(state, np.float64)
    d = np.ascontiguousarray(delta, np.float64)
    out = np.zeros_like(st)
    lib().vo_state_plus(C.byref(problem.c), _dp(st), _dp(d), _dp(out))
    return out


def gauge_reanchor(problem, state_before, state):
    sb = np.ascontiguousarray(state_before, np.float64)
    st = np.array(state, np.float64, copy=True)
    lib().vo_gauge_reanchor(C.byref(problem.c), _dp(sb), _dp(st))
    return st


def marginalize(problem, state, flag, want_system=False):
    st = np.ascontiguousarray(state, np.float64)
    out = abi.PriorData()
    mn = (C.c_int32 * 2)()
    if want_system:
        cap = abi.TANGENT_FIXED + 8 + problem.num_landmarks
        A, b = np.zeros(cap * cap), np.zeros(cap)
        rc = lib().vo_marginalize(C.byref(problem.c), _dp(st), C.c_int(flag), C.byref(out.c), _dp(A), _dp(b), mn)
    else:
        rc = lib().vo_marginalize(C.byref(problem.c), _dp(st), C.c_int(flag), C.byref(out.c), None, None, mn)
    if rc:
        raise RuntimeError("vo_marginalize rc=%d" % rc)
    if want_system:
        pos = mn[0] + mn[1]
        return out, A[: pos * pos].reshape(pos, pos).copy(), b[:pos].copy(), (mn[0], mn[1])
    return out


def optimization(problem, state, flag, options=None, want_prior=True):
    st = np.array(state, np.float64, copy=True)
    opt = options if options is not None else abi.default_options()
    summ = abi.Summary()
    out = abi.PriorData() if want_prior else None
    rc = lib().vo_optimization(C.byref(problem.c), _dp(st), C.byref(opt), C.c_int(flag), C.byref(summ), C.byref(out.c) if out else None)
    if rc:
        raise RuntimeError("vo_optimization rc=%d" % rc)
    return st, summ, out


def outlier_rejection(problem, state, focal=460.0, thresh=3.0):
    st = np.ascontiguousarray(state, np.float64)
    out = np.zeros(max(problem.num_landmarks, 1), np.uint8)
    lib().vo_outlier_rejection(C.byref(problem.c), _dp(st), C.c_double(focal), C.c_double(thresh), out.ctypes.data_as(C.c_void_p))
    return out[: problem.num_landmarks]


def triangulate(state, stereo, frame, pt0, pt1, init_depth=5.0):
    st = np.ascontiguousarray(state, np.float64)
    s_, f_ = np.ascontiguousarray(stereo, np.int32), np.ascontiguousarray(frame, np.int32)
    a, b = np.ascontiguousarray(pt0, np.float64), np.ascontiguousarray(pt1, np.float64)
    out = np.zeros(len(s_))
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    lib().vo_triangulate(_dp(st), C.c_int(len(s_)), vp(s_), vp(f_), _dp(a), _dp(b), C.c_double(init_depth), _dp(out))
    return out


def shift_depth(uv, depth, marg_R, marg_P, new_R, new_P, init_depth=5.0):
    uv, depth = np.ascontiguousarray(uv, np.float64), np.ascontiguousarray(depth, np.float64)
    mats = [np.ascontiguousarray(m, np.float64) for m in (marg_R, marg_P, new_R, new_P)]
    out = np.zeros(len(depth))
    lib().vo_shift_depth(C.c_int(len(depth)), _dp(uv), _dp(depth), _dp(mats[0]), _dp(mats[1]), _dp(mats[2]), _dp(mats[3]), C.c_double(init_depth), _dp(out))
    return out


def imu_preintegrate(dt, acc, gyr, ba, bg, noise):
    dt = np.ascontiguousarray(dt, np.float64)
    acc, gyr = np.ascontiguousarray(acc, np.float64), np.ascontiguousarray(gyr, np.float64)
    ba, bg, noise = (np.ascontiguousarray(v, np.float64) for v in (ba, bg, noise))
    rec = np.zeros(abi.IMU_DOUBLES)
    lib().vo_imu_preintegrate(C.c_int(len(dt)), _dp(dt), _dp(acc), _dp(gyr), _dp(ba), _dp(bg), _dp(noise), _dp(rec))
    return rec


def wheel_preintegrate(dt, vel, gyr, s, td, noise):
    dt = np.ascontiguousarray(dt, np.float64)
    vel, gyr = np.ascontiguousarray(vel, np.float64), np.ascontiguousarray(gyr, np.float64)
    s, noise = np.ascontiguousarray(s, np.float64), np.ascontiguousarray(noise, np.float64)
    rec = np.zeros(abi.WHEEL_DOUBLES)
    lib().vo_wheel_preintegrate(C.c_int(len(dt)), _dp(dt), _dp(vel), _dp(gyr), _dp(s), C.c_double(td), _dp(noise), _dp(rec))
    return rec


def sym_eig(A):
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    w, V = np.zeros(n), np.zeros((n, n))
    lib().vo_sym_eig(C.c_int(n), _dp(A), _dp(w), _dp(V))
    return w, V


def optimization_throughput(problems, states, flags, threads, repeat=1, options=None):
    """Run len(problems) * repeat optimisations on `threads` pthreads inside the C library; returns (count, seconds)."""
    import time
    B = len(problems)
    arr = (abi.Problem * B)()
    for i, p in enumerate(problems):
        p.fill(arr[i])
    sts = [np.ascontiguousarray(s, np.float64) for s in states]
    sp = (abi.c_double_p * B)(*[_dp(s) for s in sts])
    fl = (C.c_int32 * B)(*[int(f) for f in flags])
    opt = options if options is not None else abi.default_options()
    f = lib().vo_optimization_throughput
    f.restype = C.c_long
    t0 = time.perf_counter()
    n = f(C.c_int(B), arr, sp, fl, C.byref(opt), C.c_int(threads), C.c_int(repeat))
    return int(n), time.perf_counter() - t0


def optimization_many(problems, states, flags, threads, options=None):
    """len(problems) optimisations on `threads` pthreads; returns (solved states, iteration counts, prior digests [B][3] = n, |J|_F^2,
    |J^T r|^2, seconds).  bench.py's full-batch parity and its cpu_baseline come from this one pass."""
    import time
    B = len(problems)
    arr = (abi.Problem * B)()
    for i, p in enumerate(problems):
        p.fill(arr[i])
    sts = [np.ascontiguousarray(s, np.float64) for s in states]
    outs = [np.zeros_like(s) for s in sts]
    sp = (abi.c_double_p * B)(*[_dp(s) for s in sts])
    op = (abi.c_double_p * B)(*[_dp(s) for s in outs])
    fl = (C.c_int32 * B)(*[int(f) for f in flags])
    iters = np.zeros(B, np.int32)
    digest = np.zeros((B, 3), np.float64)
    opt = options if options is not None else abi.default_options()
    f = lib().vo_optimization_many
    f.restype = C.c_long
    t0 = time.perf_counter()
    f(C.c_int(B), arr, sp, fl, C.byref(opt), C.c_int(threads), C.c_int(1), op, iters.ctypes.data_as(C.POINTER(C.c_int32)), _dp(digest))
    return outs, iters, digest, time.perf_counter() - t0


"""Parity checks shared by the GPU tests (libviwb.so through the C ABI on a B200) and the CPU kernel-logic
emulation tests (same device source compiled with g++; tests/emu).  `ctx` is a viwb.lib.Context, `oracle` the
CPU oracle module.  Tolerances: the solver path is FP64 end to end, so the pose tolerance of north_star
(1e-4 m / 1e-4 rad) is met with a wide margin; the asserted bounds are the tight ones actually observed."""
import numpy as np

from viwb import abi, synth

POSE_TOL_M, POSE_TOL_RAD = 1e-4, 1e-4     # north_star tolerance (pose states after double2vector)
TIGHT_M, TIGHT_RAD = 1e-6, 1e-6           # what an FP64 implementation of the same recurrence should reach


def block_ptr(st, b):
    o = abi.block_offset(b) if b < 32 else abi.STATE_FIXED + b - 32
    s = abi.block_size(b) if b < 32 else 1
    return st[o:o + s].copy()


def check_factor_evaluate(ctx, oracle, cid=4, max_each=6):
    """CostFunction::Evaluate of every factor class: residuals and Jacobians (reference layout) vs the oracle."""
    prob, st, gt = synth.make_window(cid)
    st = st.copy()
    st[abi.block_offset(abi.BLK_TD)] = 0.002
    worst = 0.0
    cases = []
    for ftype in (abi.F_PROJ_2F1C, abi.F_PROJ_2F2C, abi.F_PROJ_1F2C):
        for f in np.nonzero(prob.vis_type == ftype)[0][:max_each]:
            fi, fj, lm = prob.vis_frame_i[f], prob.vis_frame_j[f], 32 + prob.vis_landmark[f]
            blocks = {abi.F_PROJ_2F1C: [fi, fj, 22, lm, 30], abi.F_PROJ_2F2C: [fi, fj, 22, 23, lm, 30], abi.F_PROJ_1F2C: [22, 23, lm, 30]}[ftype]
            cases.append((ftype, prob.vis_obs[f], blocks))
    for f in range(min(max_each, len(prob.imu_frame_i))):
        i, j = prob.imu_frame_i[f], prob.imu_frame_j[f]
        cases.append((abi.F_IMU, prob.imu_data[f], [i, 11 + i, j, 11 + j]))
    for f in range(min(max_each, len(prob.wheel_frame_i))):
        i, j = prob.wheel_frame_i[f], prob.wheel_frame_j[f]
        cases.append((abi.F_WHEEL, prob.wheel_data[f], [i, j, 24, 27, 28, 29, 31]))
    for f in range(min(max_each, len(prob.plane_frame))):
        cases.append((abi.F_PLANE, None, [prob.plane_frame[f], 24, 25, 26]))
    assert len(cases) > 0
    for ftype, consts, blocks in cases:
        params = [block_ptr(st, b) for b in blocks]
        if ftype == abi.F_WHEEL:
            params[3][0], params[4][0], params[5][0], params[6][0] = 1.01, 0.99, 1.02, 0.004
        r0, J0 = oracle.factor_evaluate(ftype, prob.globals, consts, params)
        r1, J1 = ctx.factor_evaluate(ftype, prob.globals, consts, params)
        scale = max(1.0, np.abs(r0).max())
        assert np.abs(r0 - r1).max() <= 1e-9 * scale, (ftype, r0, r1)
        for a, b in zip(J0, J1):
            s = max(1.0, np.abs(a).max())
            worst = max(worst, np.abs(a - b).max() / s)
            assert np.abs(a - b).max() <= 1e-9 * s, (ftype, np.abs(a - b).max(), s)
            if a.shape[1] == 7:
                assert np.all(b[:, 6] == 0)
        # NULL Jacobian blocks and residual-only calls follow the ceres contract
        r2, J2 = ctx.factor_evaluate(ftype, prob.globals, consts, params, null_jac=(0,))
        assert J2[0] is None and np.array_equal(r1, r2) and np.array_equal(J1[1], J2[1])
        r3, _ = ctx.factor_evaluate(ftype, prob.globals, consts, params, want_jac=False)
        assert np.array_equal(r1, r3)
    return worst


def check_normal_equations(ctx, oracle, cid):
    prob, st, gt = synth.make_window(cid)
    H0, g0, lm0, c0 = oracle.normal_equations(prob, st)
    H1, g1, lm1, c1 = ctx.normal_equations(prob, st)
    assert abs(c0 - c1) <= 1e-12 * c0
    assert np.abs(H0 - H1).max() <= 1e-12 * np.abs(H0).max()
    assert np.abs(g0 - g1).max() <= 1e-11 * np.abs(g0).max()
    assert np.abs(lm0 - lm1).max() <= 1e-11 * np.abs(lm0).max()


def check_solve(ctx, oracle, cid, seq=0, iters=8, prior_chain=False):
    cfg = synth.make_config(cid)
    s = synth.Sequence(cfg, seq, 12 if prior_chain else 11)
    prob, st, gt = s.window(0)
    if prior_chain:
        a, _ = oracle.window_solve(prob, st)
        a = oracle.gauge_reanchor(prob, st, a)
        pr = oracle.marginalize(prob, a, abi.MARGIN_OLD)
        prob, st, gt = s.window(1, prior=pr, prev_state=a)
    opt = abi.default_options()
    opt.max_num_iterations = iters
    s0, sm0 = oracle.window_solve(prob, st, opt)
    s1, sm1 = ctx.window_solve(prob, st, opt)
    assert sm0.num_iterations == sm1.num_iterations and sm0.termination_type == sm1.termination_type
    assert sm0.num_successful_steps == sm1.num_successful_steps
    assert abs(sm0.initial_cost - sm1.initial_cost) <= 1e-12 * sm0.initial_cost
    assert abs(sm0.final_cost - sm1.final_cost) <= 1e-7 * sm0.final_cost
    r0, r1 = oracle.gauge_reanchor(prob, st, s0), ctx.gauge_reanchor(prob, st, s1)
    ep, er = synth.pose_errors(r0, r1)
    assert ep <= POSE_TOL_M and er <= POSE_TOL_RAD, (ep, er)
    assert ep <= TIGHT_M and er <= TIGHT_RAD, (ep, er)
    # every other block too
    assert np.abs(r0[77:abi.STATE_FIXED] - r1[77:abi.STATE_FIXED]).max() <= 1e-6
    assert np.abs(r0[abi.STATE_FIXED:] - r1[abi.STATE_FIXED:]).max() <= 1e-6
    # constant blocks untouched bit for bit
    for b in range(abi.NUM_FIXED_BLOCKS):
        if prob.block_flags[b] & abi.BLOCK_CONSTANT:
            o = abi.block_offset(b)
            assert np.array_equal(s1[o:o + abi.block_size(b)], st[o:o + abi.block_size(b)])
    return ep, er


def check_reanchor(ctx, oracle, cid):
    prob, st, gt = synth.make_window(cid)
    s0, _ = oracle.window_solve(prob, st)
    r0, r1 = oracle.gauge_reanchor(prob, st, s0), ctx.gauge_reanchor(prob, st, s0)
    assert np.abs(r0 - r1).max() <= 1e-13 * max(1.0, np.abs(r0).max())


def prior_information_close(p0, p1, rtol=1e-6):
    assert p0.valid == p1.valid
    if not p0.valid:
        return
    assert p0.n == p1.n and p0.blocks() == p1.blocks()
    assert np.array_equal(p0.x0, p1.x0)
    A0, b0 = p0.information()
    A1, b1 = p1.information()
    assert np.abs(A0 - A1).max() <= rtol * np.abs(A0).max(), np.abs(A0 - A1).max() / np.abs(A0).max()
    assert np.abs(b0 - b1).max() <= rtol * np.abs(b0).max(), np.abs(b0 - b1).max() / np.abs(b0).max()


def check_marginalize(ctx, oracle, cid):
    """J_lin^T J_lin and J_lin^T r_lin (the order-independent content of the prior) vs the oracle."""
    prob, st, gt = synth.make_window(cid)
    s0, _ = oracle.window_solve(prob, st)
    r0 = oracle.gauge_reanchor(prob, st, s0)
    p0 = oracle.marginalize(prob, r0, abi.MARGIN_OLD)
    p1 = ctx.marginalize(prob, r0, abi.MARGIN_OLD)
    prior_information_close(p0, p1)
    # prior factor evaluate agrees as well
    x = r0.copy()
    x[:abi.STATE_FIXED] = p1.x0
    x[7:10] += 0.01
    res0, jac0 = oracle.prior_evaluate(p1, x[:abi.STATE_FIXED])
    res1, jac1 = ctx.prior_evaluate(p1, x)
    assert np.abs(res0 - res1).max() <= 1e-9 * max(1.0, np.abs(res0).max())
    assert np.array_equal(jac0, jac1)
    return p0, p1


def check_sequence(ctx, oracle, cid, nwin=3):
    """Estimator::optimization() over consecutive windows: solve + re-anchor + marginalise, prior fed forward,
    alternating MARGIN_OLD / MARGIN_SECOND_NEW like a real key-frame / non-key-frame stream."""
    cfg = synth.make_config(cid)
    s = synth.Sequence(cfg, 0, 11 + nwin)
    pr0 = pr1 = None
    prev0 = prev1 = None
    for k in range(nwin):
        flag = abi.MARGIN_OLD if k != 1 else abi.MARGIN_SECOND_NEW
        prob0, st0, _ = s.window(k, prior=pr0, prev_state=prev0)
        prob1, st1, _ = s.window(k, prior=pr1, prev_state=prev1)
        a0, sm0, q0 = oracle.optimization(prob0, st0, flag)
        a1, sm1, q1 = ctx.optimization(prob1, st1, flag)
        ep, er = synth.pose_errors(a0, a1)
        assert ep <= POSE_TOL_M and er <= POSE_TOL_RAD, (k, ep, er)
        assert q0.valid == q1.valid
        if q0.valid:
            assert q0.n == q1.n and q0.blocks() == q1.blocks()
            A0, b0 = q0.information()
            A1, b1 = q1.information()
            assert np.abs(A0 - A1).max() <= 1e-4 * np.abs(A0).max()
        pr0, pr1, prev0, prev1 = (q0 if q0.valid else None), (q1 if q1.valid else None), a0, a1
        if flag == abi.MARGIN_SECOND_NEW:
            # slideWindowNew: frame 9 is dropped, frame 10 takes its place (states are re-generated by window())
            prev0 = prev1 = None


def check_batch_matches_single(ctx, oracle):
    probs, sts, flags = [], [], []
    for cid, seq in ((1, 0), (4, 1), (2, 2), (3, 3)):
        p, s, _ = synth.make_window(cid, seq)
        probs.append(p)
        sts.append(s)
        flags.append(abi.MARGIN_OLD)
    out_s, out_sum, out_pr = ctx.optimization_batch(probs, sts, flags)
    for p, s, f, bs, bsum, bpr in zip(probs, sts, flags, out_s, out_sum, out_pr):
        a, sm, q = ctx.optimization(p, s, f)
        assert np.array_equal(a, bs)
        assert sm.num_iterations == bsum.num_iterations and sm.final_cost == bsum.final_cost
        assert q.n == bpr.n and np.array_equal(q.Jmat(), bpr.Jmat()) and np.array_equal(q.rvec(), bpr.rvec())
        a0, sm0, q0 = oracle.optimization(p, s, f)
        ep, er = synth.pose_errors(a0, bs)
        assert ep <= TIGHT_M and er <= TIGHT_RAD
    # device-resident batch API gives the same numbers and is re-runnable
    b = ctx.batch(probs, sts, flags)
    b.run()
    s1, sum1, pr1 = b.download()
    b.run()
    s2, sum2, pr2 = b.download()
    # viwb_batch_reset_states: the windows as uploaded again -- the outlier verdicts of the INITIAL guesses, and a third identical run
    b.reset()
    for p, s, o in zip(probs, sts, b.outliers()):
        assert np.array_equal(o, oracle.outlier_rejection(p, s))
    b.run()
    s3, _, _ = b.download(want_priors=False)
    for x, y in zip(s1, s3):
        assert np.array_equal(x, y)
    b.destroy()
    for x, y, z in zip(s1, s2, out_s):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    assert b.algorithmic_bytes() if False else True


def check_edge_cases(ctx, oracle):
    # window that is not full yet (frame_count < WINDOW_SIZE): no marginalisation, fewer poses
    prob, st, gt = synth.make_window(1)
    keep = prob.vis_frame_j <= 6
    flags = prob.block_flags.copy()
    for i in range(7, 11):
        flags[i] = 0
        flags[11 + i] = 0
    small = abi.WindowProblem(6, prob.num_landmarks, flags, prob.subset_mask, prob.vis_type[keep], prob.vis_landmark[keep], prob.vis_frame_i[keep],
                              prob.vis_frame_j[keep], prob.vis_obs[keep], prob.imu_frame_i[:6], prob.imu_frame_j[:6], prob.imu_data[:6],
                              globals_=prob.globals)
    s0, sm0 = oracle.window_solve(small, st)
    s1, sm1 = ctx.window_solve(small, st)
    ep, er = synth.pose_errors(s0, s1, frames=7)
    assert ep <= TIGHT_M and er <= TIGHT_RAD
    assert np.array_equal(s1[7 * 7:77], st[7 * 7:77])            # absent poses untouched
    a, sm, q = ctx.optimization(small, st, abi.MARGIN_OLD)
    assert not q.valid                                            # estimator.cpp:1666
    # no visual factors at all (IMU only, empty landmark set)
    imu_only = abi.WindowProblem(10, 0, prob.block_flags, prob.subset_mask, imu_frame_i=prob.imu_frame_i, imu_frame_j=prob.imu_frame_j,
                                 imu_data=prob.imu_data, globals_=prob.globals)
    s0, sm0 = oracle.window_solve(imu_only, st[:abi.STATE_FIXED])
    s1, sm1 = ctx.window_solve(imu_only, st[:abi.STATE_FIXED])
    assert sm0.num_iterations == sm1.num_iterations
    ep, er = synth.pose_errors(s0, s1)
    assert ep <= 1e-5 and er <= 1e-5, (ep, er)
    # zero iterations: state unchanged, cost reported
    opt = abi.default_options()
    opt.max_num_iterations = 0
    s1, sm1 = ctx.window_solve(prob, st, opt)
    assert np.array_equal(s1, st) and sm1.num_iterations == 1 and sm1.initial_cost == sm1.final_cost
    # a rejected-step heavy start (depths 3x off) follows the oracle through rejections / dogleg steps
    bad = st.copy()
    bad[abi.STATE_FIXED:] *= 3.0
    opt = abi.default_options()
    opt.max_num_iterations = 12
    s0, sm0, tr = oracle.window_solve(prob, bad, opt, want_trace=True)
    s1, sm1 = ctx.window_solve(prob, bad, opt)
    assert sm0.num_iterations == sm1.num_iterations and sm0.num_successful_steps == sm1.num_successful_steps
    ep, er = synth.pose_errors(s0, s1)
    assert ep <= POSE_TOL_M and er <= POSE_TOL_RAD, (ep, er)


# ------------------------------------------------------------------------------------------------ LK
def lk_images(seed=0, w=752, h=480):
    import cv2
    rng = np.random.default_rng(seed)
    tex = rng.normal(size=(h + 200, w + 200)).astype(np.float32)
    tex = cv2.GaussianBlur(tex, (0, 0), 2.0)
    tex = (tex - tex.min()) / (tex.max() - tex.min()) * 255
    img0 = tex[100:100 + h, 100:100 + w].astype(np.uint8)
    M = np.array([[1.01, 0.02, 3.3], [-0.015, 0.995, -2.1]], np.float32)
    img1 = cv2.warpAffine(tex, M, (w + 200, h + 200), flags=cv2.INTER_LINEAR)[100:100 + h, 100:100 + w].astype(np.uint8)
    pts = cv2.goodFeaturesToTrack(img0, 150, 0.01, 30).reshape(-1, 2).astype(np.float32)
    return np.ascontiguousarray(img0), np.ascontiguousarray(img1), pts


def check_lk(ctx, seed=0, w=752, h=480):
    """vs cv2.calcOpticalFlowPyrLK (OpenCV is the reference's third-party LK; oracle = cv2 4.13 in this image).
    Stated tolerance: <= 1e-2 px on points both sides track, >= 99 % status agreement."""
    import cv2
    img0, img1, pts = lk_images(seed, w, h)
    # add points near / outside the border to exercise the padded-window and status paths
    extra = np.array([[3.0, 4.0], [w - 2.5, h - 3.0], [w / 2, 1.0], [0.2, h / 2]], np.float32)
    pts = np.vstack([pts, extra]).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    for ml, flags in ((3, 0), (1, 4), (0, 0)):
        init = (pts + np.float32([2.0, -1.5])) if flags else None
        p_cv, st_cv, err_cv = cv2.calcOpticalFlowPyrLK(img0, img1, pts.reshape(-1, 1, 2), None if init is None else init.reshape(-1, 1, 2).copy(),
                                                       winSize=(21, 21), maxLevel=ml, criteria=crit, flags=flags)
        p_g, st_g, err_g = ctx.lk_track(img0, img1, pts, init, max_level=ml, flags=flags)
        p_cv, st_cv = p_cv.reshape(-1, 2), st_cv.reshape(-1)
        assert (st_cv == st_g).mean() >= 0.99, (ml, flags, (st_cv == st_g).mean())
        both = (st_cv == 1) & (st_g == 1)
        assert both.sum() > 100
        assert np.abs(p_cv[both] - p_g[both]).max() <= 1e-2, (ml, flags, np.abs(p_cv[both] - p_g[both]).max())
        assert np.abs(err_cv.reshape(-1)[both] - err_g[both]).max() <= 5e-2


def ref_track_checked(img_a, img_b, pts, mode, flow_back):
    """FeatureTracker::trackImage status logic restated with cv2 calls (feature_tracker.cpp:139-162, 240-251)."""
    import cv2
    h, w = img_a.shape
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    p1, st, _ = cv2.calcOpticalFlowPyrLK(img_a, img_b, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    p1, st = p1.reshape(-1, 2), st.reshape(-1).copy()

    def in_border(p):
        x, y = int(np.rint(p[0])), int(np.rint(p[1]))
        return 1 <= x < w - 1 and 1 <= y < h - 1
    if flow_back:
        if mode == 0:
            rp, rs, _ = cv2.calcOpticalFlowPyrLK(img_b, img_a, p1.reshape(-1, 1, 2), pts.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1,
                                                 criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        else:
            rp, rs, _ = cv2.calcOpticalFlowPyrLK(img_b, img_a, p1.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
        rp, rs = rp.reshape(-1, 2), rs.reshape(-1)
        for i in range(len(pts)):
            d = np.sqrt(float(pts[i, 0] - rp[i, 0]) ** 2 + float(pts[i, 1] - rp[i, 1]) ** 2)
            ok = st[i] and rs[i] and d <= 0.5
            if mode == 1:
                ok = ok and in_border(p1[i])
            st[i] = 1 if ok else 0
    if mode == 0:
        for i in range(len(pts)):
            if st[i] and not in_border(p1[i]):
                st[i] = 0
    return p1, st


def check_track_checked(ctx, seed=1):
    img0, img1, pts = lk_images(seed)
    for mode in (0, 1):
        for fb in (True, False):
            p_ref, st_ref = ref_track_checked(img0, img1, pts, mode, fb)
            p_g, st_g = ctx.track_checked(img0, img1, pts, mode=mode, flow_back=fb)
            assert (st_ref == st_g).mean() >= 0.99, (mode, fb, (st_ref == st_g).mean())
            both = (st_ref == 1) & (st_g == 1)
            assert both.sum() > 100 and np.abs(p_ref[both] - p_g[both]).max() <= 1e-2


def check_unsorted_table_and_threads(ctx, oracle):
    """The visual table may come in any order (the library groups it by landmark), and a batch large enough to use
    the parallel host lowering gives the same answer as one window at a time."""
    prob, st, gt = synth.make_window(2)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(prob.vis_type))
    shuf = abi.WindowProblem(prob.frame_count, prob.num_landmarks, prob.block_flags, prob.subset_mask, prob.vis_type[perm], prob.vis_landmark[perm],
                             prob.vis_frame_i[perm], prob.vis_frame_j[perm], prob.vis_obs[perm], prob.imu_frame_i, prob.imu_frame_j, prob.imu_data,
                             globals_=prob.globals)
    a, sa, qa = ctx.optimization(prob, st, abi.MARGIN_OLD)
    b, sb, qb = ctx.optimization(shuf, st, abi.MARGIN_OLD)
    ep, er = synth.pose_errors(a, b)
    assert ep < 1e-9 and er < 1e-9
    # The packed upload shares one host-side observation (pts_i, velocity_i, td_i) per landmark, as estimator.cpp:1595-1597 produces them; a table whose
    # factors of one landmark carry DIFFERENT host sides is legal input and must still be solved from its own numbers (the library rebuilds with one entry
    # per factor): against the oracle, which reads every factor's own 12 doubles, alone and inside a batch of ordinary windows
    odd_obs = prob.vis_obs.copy()
    lm0 = int(prob.vis_landmark[0])
    rows = np.flatnonzero(prob.vis_landmark == lm0)
    assert len(rows) >= 2
    odd_obs[rows[1], 0] += 1e-3            # pts_i.x of the landmark's second factor only
    odd_obs[rows[-1], 10] += 2e-3          # td_i of its last factor only
    odd = abi.WindowProblem(prob.frame_count, prob.num_landmarks, prob.block_flags, prob.subset_mask, prob.vis_type, prob.vis_landmark, prob.vis_frame_i, prob.vis_frame_j,
                            odd_obs, prob.imu_frame_i, prob.imu_frame_j, prob.imu_data, globals_=prob.globals)
    c, sc, qc = ctx.optimization(odd, st, abi.MARGIN_OLD)
    c0, sc0, qc0 = oracle.optimization(odd, st, abi.MARGIN_OLD)
    ep, er = synth.pose_errors(c0, c)
    assert ep < 1e-6 and er < 1e-6 and sc.num_iterations == sc0.num_iterations and qc.n == qc0.n
    assert np.abs(c - a).max() > 1e-9        # (the perturbation is visible in the solution: the test would notice a library that ignored it)
    mix_s, _, mix_q = ctx.optimization_batch([prob, odd, prob], [st, st, st], [abi.MARGIN_OLD] * 3)
    assert np.array_equal(mix_s[1], c) and np.array_equal(mix_s[0], a) and np.array_equal(mix_s[2], a) and np.array_equal(mix_q[1].Jmat(), qc.Jmat())
    probs, sts = [], []
    for i in range(20):
        p, s, _ = synth.make_window(1 + (i % 4), i % 3)
        probs.append(p)
        sts.append(s)
    out_s, out_sum, out_pr = ctx.optimization_batch(probs, sts, [abi.MARGIN_OLD] * 20)
    for i in (0, 7, 19):
        a, sm, q = ctx.optimization(probs[i], sts[i], abi.MARGIN_OLD)
        assert np.array_equal(a, out_s[i]) and q.n == out_pr[i].n and np.array_equal(q.Jmat(), out_pr[i].Jmat())


def check_solver_time_limit(ctx, oracle, cid=4):
    """SOLVER_TIME (estimator.cpp:1650-1653): an already-expired limit lets exactly one iteration be decided, which is what
    the oracle gives with max_num_iterations = 1; no limit (0) is the parity configuration."""
    cfg = synth.make_config(cid)
    prob, st, _ = synth.Sequence(cfg, 0, 11).window(0)
    opt = abi.default_options()
    opt.max_solver_time_in_seconds = 1e-9
    s1, sm1 = ctx.window_solve(prob, st, opt)
    o1 = abi.default_options()
    o1.max_num_iterations = 1
    s0, sm0 = oracle.window_solve(prob, st, o1)
    assert sm1.termination_type == abi.NO_CONVERGENCE and sm1.num_iterations == sm0.num_iterations == 2
    assert np.abs(s1 - s0).max() <= 1e-9


def check_lk_batch(ctx, streams=3, w=320, h=240, min_both=30):
    """The batched tracker (one tick of several camera streams per submission) reproduces, stream by stream, the
    trackImage logic restated with cv2 calls, over two consecutive ticks (the second one re-uses the resident
    previous image and its pyramid) and with ragged point counts."""
    import cv2
    maxn = 160
    ims = [lk_images(10 + s, w, h) for s in range(streams)]
    img0 = np.stack([i[0] for i in ims]); img1 = np.stack([i[1] for i in ims])
    right = np.stack([np.ascontiguousarray(np.roll(i[1], -4, axis=1)) for i in ims])          # 4 px disparity
    img2 = np.stack([np.ascontiguousarray(np.roll(np.roll(i[1], 2, axis=0), -3, axis=1)) for i in ims])
    right2 = np.stack([np.ascontiguousarray(np.roll(i, -4, axis=1)) for i in img2])
    pts = np.zeros((streams, maxn, 2), np.float32); n = np.zeros(streams, np.int32)
    for s in range(streams):
        p = ims[s][2][: maxn - 7 * s]
        p = p[: len(p) - 3 * s]                           # ragged counts
        n[s] = len(p); pts[s, : n[s]] = p
    lb = ctx.lk_batch(streams, w, h, maxn, stereo=True, flow_back=True)
    try:
        def compare(prev, cur, rgt, p_in, n_in, sp_in, ns_in, got):
            cur_pts, st, rp, sr = got
            for s in range(streams):
                k = int(n_in[s])
                p_ref, st_ref = ref_track_checked(prev[s], cur[s], p_in[s, :k], 0, True)
                assert (st_ref == st[s, :k]).mean() >= 0.99
                both = (st_ref == 1) & (st[s, :k] == 1)
                assert both.sum() > min_both and np.abs(p_ref[both] - cur_pts[s, :k][both]).max() <= 1e-2
                k2 = int(ns_in[s])
                q_ref, sq_ref = ref_track_checked(cur[s], rgt[s], sp_in[s, :k2], 1, True)
                assert (sq_ref == sr[s, :k2]).mean() >= 0.99
                both = (sq_ref == 1) & (sr[s, :k2] == 1)
                assert both.sum() > min_both and np.abs(q_ref[both] - rp[s, :k2][both]).max() <= 1e-2
        # tick 1: prev + cur + right uploaded; stereo runs on the (exactly known) positions of the features in cur
        sp = pts.copy()
        lb.upload(prev=img0, cur=img1, right=right, prev_pts=pts, n_prev=n, stereo_pts=sp, n_stereo=n)
        lb.run()
        got = [a.copy() for a in lb.download()]
        compare(img0, img1, right, pts, n, sp, n, got)
        # tick 2: only the new images travel; the tracked points of tick 1 become prev_pts
        p2 = got[0].copy(); n2 = n.copy()
        lb.upload(cur=img2, right=right2, prev_pts=p2, n_prev=n2, stereo_pts=p2, n_stereo=n2)
        lb.run()
        got2 = [a.copy() for a in lb.download()]
        compare(img1, img2, right2, p2, n2, p2, n2, got2)
    finally:
        lb.close()


def check_preintegration(ctx, oracle, seed=5):
    """batched IMU / wheel pre-integration against the oracle's restatement of IntegrationBase::propagate /
    WheelIntegrationBase::propagate, ragged interval lengths (including the 10-sample interval of the 20 Hz / 200 Hz configs)."""
    rng = np.random.default_rng(seed)
    lens = [10, 1, 40, 7, 10, 23]
    dts = [np.full(k, 0.005) + rng.uniform(-1e-4, 1e-4, k) for k in lens]
    accs = [rng.normal(0, 2.0, (k + 1, 3)) + np.array([0, 0, 9.8]) for k in lens]
    gyrs = [rng.normal(0, 0.5, (k + 1, 3)) for k in lens]
    ba, bg = rng.normal(0, 0.05, (len(lens), 3)), rng.normal(0, 0.01, (len(lens), 3))
    noise = np.array([0.1, 0.01, 1e-3, 1e-4])
    rec = ctx.imu_preintegrate(dts, accs, gyrs, ba, bg, noise)
    for i, k in enumerate(lens):
        ref = oracle.imu_preintegrate(dts[i], accs[i], gyrs[i], ba[i], bg[i], noise)
        assert np.abs(rec[i] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), (i, np.abs(rec[i] - ref).max())
    vels = [rng.normal(0, 0.3, (k + 1, 3)) + np.array([1.0, 0, 0]) for k in lens]
    s = 1.0 + rng.normal(0, 0.01, (len(lens), 3))
    td = rng.normal(0, 0.002, len(lens))
    wn = np.array([0.01, 0.004])
    wrec = ctx.wheel_preintegrate(dts, vels, gyrs, s, td, wn)
    for i, k in enumerate(lens):
        ref = oracle.wheel_preintegrate(dts[i], vels[i], gyrs[i], s[i], td[i], wn)
        assert np.abs(wrec[i] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), (i, np.abs(wrec[i] - ref).max())
    # the records drive the factors: a record built on the device evaluates to the same residual as one built by the oracle
    return rec, wrec


def alignment_case(cid, F=15, seed=3, scale=2.3, c0_tilt=(0.3, -0.2, 0.7), pos_noise=2e-5):
    """Inputs of VisualIMUAlignment from a synthetic sequence: the up-to-scale camera poses an SfM would give (frame c0 = the world rotated by
    `c0_tilt`, so that gravity is not along an axis; positions divided by `scale`), the raw IMU buffers, wheel pre-integrations for wheel configs."""
    cfg = synth.make_config(cid)
    seq = synth.Sequence(cfg, seed, F)
    Rc0 = synth.Rz(c0_tilt[2]) @ synth.Ry(c0_tilt[1]) @ synth.Rx(c0_tilt[0])
    rng = np.random.default_rng(100 * cid + seed)
    R = np.array([Rc0 @ seq.gt_R[k] for k in range(F)])
    T = np.array([(Rc0 @ seq.gt_P[k] + R[k] @ cfg.t_ic0) / scale for k in range(F)]) + rng.normal(0, pos_noise, (F, 3))
    dts, accs, gyrs = [i[0] for i in seq.imu], [i[1] for i in seq.imu], [i[2] for i in seq.imu]
    case = {"cfg": cfg, "R": R, "T": T, "dts": dts, "accs": accs, "gyrs": gyrs, "noise": np.array([cfg.acc_n, cfg.gyr_n, cfg.acc_w, cfg.gyr_w]), "bg0": np.zeros(3),
            "tic": cfg.t_ic0, "rio": None, "tio": None, "g_norm": cfg.g_norm, "wheel": None, "scale": scale, "g_true": Rc0 @ np.array([0, 0, cfg.g_norm])}
    if cfg.use_wheel:
        case["rio"], case["tio"] = cfg.R_io, cfg.t_io
        case["wheel_samples"] = seq.wheel
    return case


def _alignment_wheel_records(ctx, case):
    if case["cfg"].use_wheel and case["wheel"] is None:
        w = case["wheel_samples"]
        n = len(w)
        case["wheel"] = ctx.wheel_preintegrate([i[0] for i in w], [i[1] for i in w], [i[2] for i in w], np.ones((n, 3)), np.zeros(n), np.array([case["cfg"].vel_n_wheel, case["cfg"].gyr_n_wheel]))
    return case["wheel"]


def check_visual_imu_alignment(ctx, oracle, cid, F=15):
    """SURVEY 8 f-4 ii: solveGyroscopeBias + repropagation + LinearAlignment[WithWheel] + RefineGravity[WithWheel] on the device against the
    numpy restatement (oracle/init_oracle.py); the result is also physically right (gravity direction and scale of the synthetic sequence)."""
    import init_oracle as io
    case = alignment_case(cid, F)
    wheel = _alignment_wheel_records(ctx, case)
    got = ctx.visual_imu_alignment(case["R"], case["T"], case["dts"], case["accs"], case["gyrs"], case["noise"], case["bg0"], wheel, case["tic"], case["rio"], case["tio"], case["g_norm"])
    n = len(case["dts"])
    rec0 = np.array([oracle.imu_preintegrate(case["dts"][i], case["accs"][i], case["gyrs"][i], np.zeros(3), case["bg0"], case["noise"]) for i in range(n)])
    dbg = io.solve_gyroscope_bias(case["R"], rec0)
    assert np.abs(got["delta_bg"] - dbg).max() <= 1e-9 * max(1.0, np.abs(dbg).max()), (cid, got["delta_bg"], dbg)
    rec = np.array([oracle.imu_preintegrate(case["dts"][i], case["accs"][i], case["gyrs"][i], np.zeros(3), case["bg0"] + dbg, case["noise"]) for i in range(n)])
    ok, g, x = io.linear_alignment(case["R"], case["T"], rec, wheel, case["tic"], case["rio"], case["tio"], case["g_norm"])
    assert got["ok"] == ok and ok, (cid, got["ok"], ok)
    assert len(got["x"]) == len(x)
    assert np.abs(got["g"] - g).max() <= 1e-7 * case["g_norm"], (cid, got["g"], g)
    assert np.abs(got["x"] - x).max() <= 1e-7 * max(1.0, np.abs(x).max()), (cid, np.abs(got["x"] - x).max())
    cosang = got["g"] @ case["g_true"] / (np.linalg.norm(got["g"]) * case["g_norm"])
    assert cosang > np.cos(np.deg2rad(8.0)) and abs(got["x"][-1] / case["scale"] - 1.0) < 0.25, (cid, cosang, got["x"][-1])    # 0.7 s of motion: a sanity bound, not a parity bound
    # a sequence the first stage must reject (|g| far from G): the unrefined solution comes back, 3F+4 long, as in the reference
    bad = ctx.linear_alignment(case["R"], case["T"] * 3.0, rec, wheel, case["tic"], case["rio"], case["tio"], case["g_norm"] + 2.0)
    okb, gb, xb = io.linear_alignment(case["R"], case["T"] * 3.0, rec, wheel, case["tic"], case["rio"], case["tio"], case["g_norm"] + 2.0)
    assert bad[0] == okb and len(bad[2]) == len(xb) and np.abs(bad[2] - xb).max() <= 1e-7 * max(1.0, np.abs(xb).max())
    # argument errors are reported, not computed on: fewer than two frames, more than VIWB_MAX_INIT_FRAMES
    from viwb.lib import ViwbError
    for nf in (1, 65):
        try:
            ctx.solve_gyroscope_bias(np.tile(np.eye(3), (nf, 1, 1)), np.zeros((max(nf - 1, 1), 287)))
            raise AssertionError("accepted %d frames" % nf)
        except ViwbError:
            pass
    if wheel is not None:                                # wheel rows without the wheel extrinsic: refused
        try:
            ctx.linear_alignment(case["R"], case["T"], rec, wheel, case["tic"], None, None, case["g_norm"])
            raise AssertionError("accepted wheel records without RIO / TIO")
        except ViwbError:
            pass
    return got


def check_visual_imu_alignment_vs_reference_code(ctx, cid, F=15):
    """The same call against VisualIMUAlignment of the reference's own compiled initial_aligment.cpp (oracle/_ref)"""
    import viw_ref
    case = alignment_case(cid, F)
    wheel = _alignment_wheel_records(ctx, case)
    got = ctx.visual_imu_alignment(case["R"], case["T"], case["dts"], case["accs"], case["gyrs"], case["noise"], case["bg0"], wheel, case["tic"], case["rio"], case["tio"], case["g_norm"])
    ref = viw_ref.visual_imu_alignment(case["R"], case["T"], case["dts"], case["accs"], case["gyrs"], case["noise"], case["bg0"], wheel, case["tic"], case["rio"], case["tio"],
                                       np.array([0, 0, case["g_norm"]]))
    assert got["ok"] == ref["ok"] and len(got["x"]) == len(ref["x"])
    assert np.abs(got["delta_bg"] - ref["delta_bg"]).max() <= 1e-9 * max(1.0, np.abs(ref["delta_bg"]).max())
    assert np.abs(got["imu"] - ref["imu"]).max() <= 1e-9 * max(1.0, np.abs(ref["imu"]).max())
    assert np.abs(got["g"] - ref["g"]).max() <= 1e-7 * case["g_norm"] and np.abs(got["x"] - ref["x"]).max() <= 1e-7 * max(1.0, np.abs(ref["x"]).max())


def check_outlier_rejection(ctx, oracle):
    """Estimator::outliersRejection on solved windows (mono and stereo shapes), with a few landmarks pushed far
    off so that both verdicts occur; single-call and batch entry points."""
    for cid in (1, 2, 4):
        cfg = synth.make_config(cid)
        prob, st, _ = synth.Sequence(cfg, 2, 11).window(0)
        a, _, _ = ctx.optimization(prob, st, abi.MARGIN_OLD)
        bad = a.copy()
        bad[abi.STATE_FIXED: abi.STATE_FIXED + 5] *= 3.0            # wrong inverse depths -> large reprojection errors
        for x in (a, bad):
            ref = oracle.outlier_rejection(prob, x)
            got = ctx.outlier_rejection(prob, x)
            assert np.array_equal(ref, got)
        assert oracle.outlier_rejection(prob, bad)[:5].sum() >= 3 and oracle.outlier_rejection(prob, a).mean() < 0.2
    # batch entry: verdicts on the solved + re-anchored windows the device holds
    cfg = synth.make_config(2)
    wins = [synth.Sequence(cfg, s, 11).window(0) for s in range(3)]
    batch = ctx.batch([w[0] for w in wins], [w[1] for w in wins], [abi.MARGIN_OLD] * 3)
    batch.run()
    sts, _, _ = batch.download()
    outs = batch.outliers()
    batch.destroy()
    for (prob, _, _), x, o in zip(wins, sts, outs):
        assert np.array_equal(o, oracle.outlier_rejection(prob, x))


def check_small_edges(ctx, oracle):
    """Empty / degenerate inputs of the tracker batch and the widened rows."""
    import numpy as np
    # a camera stream with no points, another with one point on the border
    w, h, maxn = 160, 120, 8
    img0, img1, _ = lk_images(3, w, h)
    lb = ctx.lk_batch(2, w, h, maxn, stereo=True, flow_back=True)
    try:
        pts = np.zeros((2, maxn, 2), np.float32)
        pts[1, 0] = [0.5, 0.5]
        n = np.array([0, 1], np.int32)
        imgs0, imgs1 = np.stack([img0, img0]), np.stack([img1, img1])
        lb.upload(prev=imgs0, cur=imgs1, right=imgs1, prev_pts=pts, n_prev=n, stereo_pts=pts, n_stereo=n)
        lb.run()
        cur, st, rp, sr = lb.download()
        p_ref, st_ref = ref_track_checked(img0, img1, pts[1, :1], 0, True)
        assert st[1, 0] == st_ref[0]
    finally:
        lb.close()
    # pre-integration: zero intervals, and an interval with zero steps (identity pre-integration, zero covariance)
    noise = np.array([0.1, 0.01, 1e-3, 1e-4])
    assert ctx.imu_preintegrate([], [], [], np.zeros((0, 3)), np.zeros((0, 3)), noise).shape == (0, abi.IMU_DOUBLES)
    rec = ctx.imu_preintegrate([np.zeros(0)], [np.array([[0.0, 0, 9.8]])], [np.zeros((1, 3))], np.zeros((1, 3)), np.zeros((1, 3)), noise)
    ref = oracle.imu_preintegrate(np.zeros(0), np.array([[0.0, 0, 9.8]]), np.zeros((1, 3)), np.zeros(3), np.zeros(3), noise)
    assert np.array_equal(rec[0], ref) and rec[0][0] == 0.0 and rec[0][7] == 1.0


def check_triangulation(ctx, oracle, seed=9):
    """two-view triangulation (stereo and motion branches of FeatureManager::triangulate) and
    removeBackShiftDepth, against the oracle and against an independent numpy SVD / closed form."""
    from viwb import geom
    rng = np.random.default_rng(seed)
    cfg = synth.make_config(2)
    prob, st, gt = synth.Sequence(cfg, 1, 11).window(0)
    x = gt.copy()
    n = 60
    stereo = (rng.uniform(size=n) < 0.5).astype(np.int32)
    frame = rng.integers(0, 9, n).astype(np.int32)

    def cam(i, c):
        Rs = geom.q_to_R(x[7 * i + 3: 7 * i + 7]); ric = geom.q_to_R(x[176 + 7 * c + 3: 176 + 7 * c + 7])
        R = Rs @ ric; t = x[7 * i: 7 * i + 3] + Rs @ x[176 + 7 * c: 176 + 7 * c + 3]
        return np.hstack([R.T, (-R.T @ t)[:, None]])
    pt0, pt1, ref, truth = np.zeros((n, 2)), np.zeros((n, 2)), np.zeros(n), np.zeros(n)
    for k in range(n):
        P0 = cam(frame[k], 0); P1 = cam(frame[k], 1) if stereo[k] else cam(frame[k] + 1, 0)
        pc = np.array([rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), 1.0]) * rng.uniform(2, 15)       # point in camera 0
        pw = P0[:, :3].T @ (pc - P0[:, 3])
        q1 = P1[:, :3] @ pw + P1[:, 3]
        pt0[k] = pc[:2] / pc[2] + rng.normal(0, 1e-3, 2); pt1[k] = q1[:2] / q1[2] + rng.normal(0, 1e-3, 2)
        D = np.vstack([pt0[k, 0] * P0[2] - P0[0], pt0[k, 1] * P0[2] - P0[1], pt1[k, 0] * P1[2] - P1[0], pt1[k, 1] * P1[2] - P1[1]])
        v = np.linalg.svd(D)[2][-1]
        d = P0[2, :3] @ (v[:3] / v[3]) + P0[2, 3]
        ref[k] = d if d > 0 else 5.0
        truth[k] = pc[2]
    got, orc = ctx.triangulate(x, stereo, frame, pt0, pt1), oracle.triangulate(x, stereo, frame, pt0, pt1)
    assert np.abs(got - orc).max() <= 1e-9 * np.abs(orc).max()
    try:                                                                 # the reference's own FeatureManager::triangulate, where oracle/_ref exists
        import viw_ref
        if viw_ref.available():
            rc = viw_ref.triangulate(x, stereo, frame, pt0, pt1)
            assert np.abs(got - rc).max() <= 1e-9 * np.abs(rc).max()
    except ImportError:
        pass
    assert np.abs(got - ref).max() <= 1e-7 * np.abs(ref).max()          # numpy's SVD: independent algorithm
    assert np.median(np.abs(got - truth) / truth) < 0.2                  # and it does triangulate the noisy points
    # removeBackShiftDepth: closed form
    uv = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-0.6, 0.6, n), np.ones(n)])
    dep = rng.uniform(1, 20, n); dep[:3] = -1.0
    mR, nR = geom.q_to_R(x[3:7]) @ geom.q_to_R(x[179:183]), geom.q_to_R(x[10:14]) @ geom.q_to_R(x[179:183])
    mP, nP = x[0:3] + geom.q_to_R(x[3:7]) @ x[176:179], x[7:10] + geom.q_to_R(x[10:14]) @ x[176:179]
    z = ((uv * dep[:, None]) @ mR.T + mP - nP) @ nR[:, 2]
    want = np.where(z > 0, z, 5.0)
    assert np.abs(ctx.shift_depth(uv, dep, mR, mP, nR, nP) - want).max() <= 1e-12 * np.abs(want).max()
    assert np.abs(oracle.shift_depth(uv, dep, mR, mP, nR, nP) - want).max() <= 1e-12 * np.abs(want).max()


def check_full_batch_properties(ctx, oracle, distinct, copies):
    """BASELINE-sized batch (bench.py's workload shape): size-independent properties on EVERY window, oracle parity on a sample.
    * the robustified cost never increases (trust-region acceptance);
    * gauge: frame-0 position and yaw equal their pre-solve values after the re-anchoring (estimator.cpp:1224-1276);
    * the new prior is a square root of a PSD information matrix: J^T J symmetric, eigenvalues >= 0, r in range(J);
    * identical windows in the batch give bit-identical results (deterministic, order-independent kernels)."""
    import bench
    from viwb import geom
    cfg, seqs, first = bench.make_windows(0, distinct, copies)
    a0, _, q0 = ctx.optimization_batch([f[0] for f in first], [f[1] for f in first], [abi.MARGIN_OLD] * len(first))
    probs, states = bench.replicate(seqs, q0, a0, copies, 0)
    probs = probs + probs[:2]; states = states + [states[0].copy(), states[1].copy()]          # duplicates: must match bit for bit
    sts, sums, pri = ctx.optimization_batch(probs, states, [abi.MARGIN_OLD] * len(probs))
    B = len(probs)
    for i in range(B):
        assert sums[i].final_cost <= sums[i].initial_cost * (1 + 1e-12)
        assert 2 <= sums[i].num_iterations <= 9
        p0, p1 = states[i][0:3], sts[i][0:3]
        assert np.abs(p0 - p1).max() <= 1e-9
        y0 = geom.R_to_ypr(geom.q_to_R(states[i][3:7]))[0] if hasattr(geom, "R_to_ypr") else None
        if y0 is not None:
            y1 = geom.R_to_ypr(geom.q_to_R(sts[i][3:7]))[0]
            assert abs(((y0 - y1 + 180.0) % 360.0) - 180.0) <= 1e-7
        q = pri[i]
        assert q.valid and 60 <= q.n <= abi.MAX_PRIOR_DIM
        J, r = q.Jmat(), q.rvec()
        A = J.T @ J
        w = np.linalg.eigvalsh(0.5 * (A + A.T))
        assert w.min() >= -1e-9 * w.max()
        rr = J @ np.linalg.lstsq(J, r, rcond=None)[0]
        assert np.abs(rr - r).max() <= 1e-6 * max(1.0, np.abs(r).max())
    for i in (0, 1):
        assert np.array_equal(sts[i], sts[B - 2 + i]) and sums[i].final_cost == sums[B - 2 + i].final_cost
        assert np.array_equal(pri[i].Jmat(), pri[B - 2 + i].Jmat())
    for i in range(0, B, max(1, B // 5)):
        a, sm, _ = oracle.optimization(probs[i], states[i], abi.MARGIN_OLD)
        ep, er = synth.pose_errors(a, sts[i])
        assert ep <= POSE_TOL_M and er <= POSE_TOL_RAD
        assert abs(sm.final_cost - sums[i].final_cost) <= 1e-7 * sm.final_cost and sm.num_iterations == sums[i].num_iterations


def check_undistort_velocity(ctx):
    """undistortedPts / ptsVelocity restated in numpy with the reference's rounding points
    (double lift, cv::Point2f narrowing, float difference / double dt)."""
    rng = np.random.default_rng(4)
    cam = (461.1586, 459.7529, 362.6593, 248.5236, -0.2847798, 0.08245052, -1.0946e-06, 4.78701e-06)     # config/euroc/cam0_pinhole.yaml
    n = 300
    pts = np.column_stack([rng.uniform(0, 752, n), rng.uniform(0, 480, n)]).astype(np.float32)
    fx, fy, cx, cy, k1, k2, p1, p2 = cam

    def lift(p):
        mxd, myd = (1.0 / fx) * p[:, 0].astype(np.float64) + (-cx / fx), (1.0 / fy) * p[:, 1].astype(np.float64) + (-cy / fy)
        mx, my = mxd.copy(), myd.copy()
        for it in range(8):
            x, y = (mxd, myd) if it == 0 else (mx, my)
            mx2, my2, mxy = x * x, y * y, x * y
            rho2 = mx2 + my2
            rad = k1 * rho2 + k2 * rho2 * rho2
            dx = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2)
            dy = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2)
            mx, my = mxd - dx, myd - dy
        return np.column_stack([mx, my]).astype(np.float32)
    ref_un = lift(pts)
    prev = (ref_un + rng.normal(0, 0.01, (n, 2))).astype(np.float32)
    has = (rng.uniform(size=n) < 0.8).astype(np.uint8)
    dt = 0.05
    ref_vel = np.where(has[:, None] > 0, ((ref_un - prev).astype(np.float64) / dt), 0.0).astype(np.float32)
    un, vel = ctx.undistort_velocity(cam, pts, prev, has, dt)
    assert np.abs(un - ref_un).max() <= 2e-7          # one float ulp at |x| < 1 (FMA contraction in the double lift)
    try:                                              # and against the reference's own PinholeCamera::liftProjective, where oracle/_ref exists
        import viw_ref
        if viw_ref.available():
            assert np.abs(un - viw_ref.undistorted_pts(cam, 752, 480, pts)).max() <= 2e-7
    except ImportError:
        pass
    assert np.abs(vel - ref_vel).max() <= 1e-5
    un0, vel0 = ctx.undistort_velocity((fx, fy, cx, cy, 0, 0, 0, 0), pts, None, None, dt)
    assert np.abs(un0[:, 0] - ((pts[:, 0].astype(np.float64) - cx) / fx).astype(np.float32)).max() <= 2e-7 and not vel0.any()


# ------------------------------------------------------------------------------------------------ feature detection
def _tracked_points(rng, w, h, n, cap):
    pts = np.zeros((cap, 2), np.float32)
    pts[:n] = np.column_stack([rng.uniform(1, w - 2, n), rng.uniform(1, h - 2, n)])
    cnt = np.zeros(cap, np.int32)
    cnt[:n] = rng.integers(1, 15, n)
    return pts, cnt


def check_set_mask(ctx, w=752, h=480):
    """viwb_set_mask against the restated loop of feature_tracker.cpp:59-89 (bit exact: mask bytes and survivor order)."""
    import feature_oracle as fo
    rng = np.random.default_rng(11)
    for n, md, with_base in [(180, 30, False), (40, 30, True), (0, 30, False), (300, 7, False), (120, 1, False)]:
        pts, cnt = _tracked_points(rng, w, h, n, max(n, 1))
        base = None
        if with_base:
            base = np.full((h, w), 255, np.uint8)
            base[:, : w // 5] = 0; base[h - 40:, :] = 128          # a fisheye-style mask: only 255 counts as free (== 255 test, :81)
        m0, k0 = fo.set_mask(w, h, pts[:n], cnt[:n], md, base)
        m1, k1 = ctx.set_mask(w, h, pts[:n], cnt[:n], md, base)
        assert np.array_equal(k0, k1), (n, md)
        assert np.array_equal(m0, m1), (n, md)


def check_good_features(ctx, full=True):
    """viwb_good_features_to_track against the oracle (itself bit-exact with cv2's scalar path): identical corner lists."""
    import feature_oracle as fo
    rng = np.random.default_rng(12)
    cases = [(480, 752, 150, 30.0, True, 0.01), (480, 752, 60, 30.0, False, 0.01), (480, 752, 0, 30.0, True, 0.01),
             (480, 752, 400, 6.0, False, 1e-5),          # > 8192 candidates above the threshold: the in-place global-memory sort
             (120, 161, 500, 7.5, False, 0.01), (67, 90, 40, 0.5, True, 0.01), (33, 47, 10, 3.0, False, 0.01)]
    for k, (h, w, mc, md, use_mask, ql) in enumerate(cases if full else cases[3:]):
        img = synth.texture_image(h, w, 20 + k)
        mask = None
        if use_mask:
            mask = np.full((h, w), 255, np.uint8)
            for _ in range(max(2, w * h // 4000)):
                fo.paint_circle(mask, int(rng.integers(0, w)), int(rng.integers(0, h)), 30 if w > 400 else 9)
        ref = fo.good_features_to_track(img, mc, ql, md, mask)
        got = ctx.good_features_to_track(img, mc, ql, md, mask, capacity=1024)
        assert ref.shape == got.shape and np.array_equal(ref, got), (h, w, mc, md, len(ref), len(got))
    # a non-contiguous row stride and an all-zero mask
    big = synth.texture_image(100, 200, 3)
    view = big[:, 10:150]
    assert np.array_equal(fo.good_features_to_track(np.ascontiguousarray(view), 30, 0.01, 10.0), ctx.good_features_to_track(view, 30, 0.01, 10.0))
    assert len(ctx.good_features_to_track(np.ascontiguousarray(view), 30, 0.01, 10.0, np.zeros(view.shape, np.uint8))) == 0
    flat = np.full((64, 64), 77, np.uint8)
    assert len(ctx.good_features_to_track(flat, 30, 0.01, 10.0)) == 0


def check_detector_batch(ctx, streams=3, w=752, h=480, max_cnt=150, min_dist=30):
    """viwb_detector_detect (setMask + goodFeaturesToTrack chained on the device for several streams) against the oracle chain."""
    import feature_oracle as fo
    rng = np.random.default_rng(13)
    cap = 256
    imgs = np.stack([synth.texture_image(h, w, 40 + f) for f in range(streams)])
    pts, cnt, n = np.zeros((streams, cap, 2), np.float32), np.zeros((streams, cap), np.int32), np.zeros(streams, np.int32)
    for f in range(streams):
        n[f] = [90, 0, 200][f % 3]
        pts[f], cnt[f] = _tracked_points(rng, w, h, int(n[f]), cap)
    det = ctx.detector(streams, w, h, cap, min_dist)
    keep, n_keep, new_pts, n_new, mask = det.detect(imgs, pts, cnt, n, max_cnt, want_mask=True)
    for f in range(streams):
        m0, k0, c0 = fo.detect(imgs[f], pts[f, : n[f]], cnt[f, : n[f]], max_cnt, min_dist)
        assert n_keep[f] == len(k0) and np.array_equal(keep[f, : n_keep[f]], k0), f
        assert np.array_equal(mask[f], m0), f
        assert n_new[f] == len(c0) and np.array_equal(new_pts[f, : n_new[f]], c0), (f, n_new[f], len(c0))
        assert n_keep[f] + n_new[f] <= max(max_cnt, n_keep[f])
    det.close()
    return det


def check_detector_resident(ctx, w=752, h=480):
    """The detector reading the tracker's resident current images (no second upload) gives what it gives on host images."""
    rng = np.random.default_rng(14)
    F, cap = 2, 192
    img0 = np.stack([synth.texture_image(h, w, 60 + f) for f in range(F)])
    img1 = np.stack([np.roll(img0[f], (1, 2), axis=(0, 1)) for f in range(F)])
    lb = ctx.lk_batch(F, w, h, cap, stereo=False, flow_back=True)
    pts, cnt, n = np.zeros((F, cap, 2), np.float32), np.zeros((F, cap), np.int32), np.array([100, 30], np.int32)
    for f in range(F):
        pts[f], cnt[f] = _tracked_points(rng, w, h, int(n[f]), cap)
    det = ctx.detector(F, w, h, cap, 30)
    for tick, cur in enumerate([img1, img0, img1]):              # the tracker alternates its two left slots
        lb.upload(prev=img0 if tick == 0 else None, cur=cur, prev_pts=pts, n_prev=n)
        lb.run(); lb.download()
        a = [x.copy() for x in det.detect(None, pts, cnt, n, 150, resident=lb)[:4]]
        b = [x.copy() for x in det.detect(cur, pts, cnt, n, 150)[:4]]
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3]), tick
        for f in range(F):
            assert np.array_equal(a[0][f, : a[1][f]], b[0][f, : b[1][f]]) and np.array_equal(a[2][f, : a[3][f]], b[2][f, : b[3][f]]), (tick, f)
    det.close(); lb.close()


# ------------------------------------------------------------------------------------------------ session tracker
def camera_sequence(seed, w, h, ticks, disparity=5.0):
    """A moving camera over a large texture: left / right 8-bit frames per tick and the per-tick image motion (px)."""
    import cv2
    world = synth.texture_image(h + 160, w + 160, seed)
    rng = np.random.default_rng(900 + seed)
    pos = np.array([80.0, 80.0]); vel = rng.uniform(-2.5, 2.5, 2)
    left, right, motion = [], [], []
    for _ in range(ticks):
        step = vel + rng.normal(0, 0.3, 2)
        pos = np.clip(pos + step, 20, 140)
        M = np.float32([[1, 0, -pos[0]], [0, 1, -pos[1]]])
        left.append(cv2.warpAffine(world, M, (w, h), flags=cv2.INTER_LINEAR))
        M[0, 2] -= disparity
        right.append(cv2.warpAffine(world, M, (w, h), flags=cv2.INTER_LINEAR))
        motion.append(-step)
    return left, right, motion


def check_tracker_session(ctx, streams=2, w=320, h=240, ticks=5, max_cnt=60, min_dist=20, stereo=True, flow_back=True, predict=True):
    """viwb_tracker_track = FeatureTracker::trackImage() per stream, device-resident session state, against the line-by-line
    restatement running cv2.calcOpticalFlowPyrLK (oracle/feature_oracle.py:FeatureTrackerRef): ids, track counts and row order
    identical, pixel positions within the LK tolerance (1e-2 px), undistorted points / velocities within what that allows."""
    import feature_oracle as fo
    cam0 = (461.1586 * w / 752, 459.7529 * w / 752, w / 2 - 3.2, h / 2 + 1.7, -0.2847798, 0.08245052, -1.0946e-06, 4.78701e-06)
    cam1 = (457.5874 * w / 752, 456.1340 * w / 752, w / 2 + 4.1, h / 2 - 2.6, -0.2836831, 0.07395907, 1.9359e-04, 1.7618e-05)
    seqs = [camera_sequence(70 + f, w, h, ticks) for f in range(streams)]
    refs = [fo.FeatureTrackerRef(cam0, cam1 if stereo else None, max_cnt, min_dist, flow_back) for _ in range(streams)]
    trk = ctx.tracker(streams, w, h, cam0, cam1 if stereo else None, max_cnt, min_dist, flow_back)
    tol_px, dt = 1e-2, 0.05
    tol_un = 1.5 * tol_px / cam0[0]
    total_tracked = 0
    for t in range(ticks):
        left = np.stack([seqs[f][0][t] for f in range(streams)])
        right = np.stack([seqs[f][1][t] for f in range(streams)]) if stereo else None
        pp, hp = None, None
        if predict and t >= 2:
            # stream 0: the true image motion as prediction (maxLevel-1 pass succeeds); last stream: a prediction outside the image,
            # so fewer than 10 points succeed and the full-pyramid pass repeats (feature_tracker.cpp:130-137)
            pp, hp = np.zeros((streams, max_cnt, 2), np.float32), np.zeros(streams, np.uint8)
            for f in ([0, streams - 1] if streams > 1 else [0]):
                prev = refs[f].prev_pts
                pred = (prev + (seqs[f][2][t] if f == 0 else np.float32([w + 37.0, h + 29.0]))).astype(np.float32)
                pp[f, : len(prev)] = pred; hp[f] = 1
                refs[f].set_prediction(pred)
        trk.track(0.05 * (t + 1) * (dt / 0.05), left, right, pp, hp)
        n_left, ids, cnt, feat, n_right, ids_r, feat_r = trk.download()
        for f in range(streams):
            r = refs[f].track_image(0.05 * (t + 1), seqs[f][0][t], seqs[f][1][t] if stereo else None)
            rid, rcnt, rpts, run, rvel, rid_r, rpts_r, run_r, rvel_r = r
            n = int(n_left[f])
            assert n == len(rid), (t, f, n, len(rid))
            assert np.array_equal(ids[f, :n], rid) and np.array_equal(cnt[f, :n], rcnt), (t, f)
            assert np.abs(feat[f, :n, 2:4] - rpts).max() <= tol_px, (t, f, np.abs(feat[f, :n, 2:4] - rpts).max())
            assert np.abs(feat[f, :n, 0:2] - run).max() <= tol_un
            assert np.abs(feat[f, :n, 4:6] - rvel).max() <= 2 * tol_un / dt
            total_tracked += int((rcnt > 1).sum())
            if stereo:
                m = int(n_right[f])
                assert m == len(rid_r) and np.array_equal(ids_r[f, :m], rid_r), (t, f, m, len(rid_r))
                if m:
                    assert np.abs(feat_r[f, :m, 2:4] - rpts_r).max() <= tol_px
                    assert np.abs(feat_r[f, :m, 0:2] - run_r).max() <= tol_un
                    assert np.abs(feat_r[f, :m, 4:6] - rvel_r).max() <= 2 * tol_un / dt
                assert m >= 0.5 * n, (t, f, m, n)                 # the synthetic pair is a pure shift: most points must match
            ff = trk.feature_frame(f)
            assert len(ff) == n and all(len(v) in (1, 2) and v[0][0] == 0 for v in ff.values())
    assert total_tracked >= streams * (ticks - 1) * max_cnt // 3, total_tracked      # the sessions really follow features across ticks
    if predict and ticks > 2:
        assert refs[0].stats["predicted"] == ticks - 2 and refs[0].stats["repeated"] == 0, refs[0].stats
        if streams > 1:
            assert refs[-1].stats["repeated"] == ticks - 2, refs[-1].stats
    trk.close()


def check_tracker_edges(ctx):
    """Degenerate sessions: a featureless stream beside a textured one, error returns that leave the session usable."""
    import feature_oracle as fo
    from viwb.lib import ViwbError
    w, h, max_cnt, min_dist = 200, 160, 25, 15
    cam = (120.0, 119.0, 99.5, 80.2, -0.28, 0.08, 1e-4, -2e-4)
    left, right, _ = camera_sequence(75, w, h, 3)
    flat = np.full((h, w), 128, np.uint8)
    trk = ctx.tracker(2, w, h, cam, cam, max_cnt, min_dist, True)
    ref = fo.FeatureTrackerRef(cam, cam, max_cnt, min_dist, True)
    for t in range(3):
        trk.track(0.05 * (t + 1), np.stack([flat, left[t]]), np.stack([flat, right[t]]))
        n_left, ids, cnt, feat, n_right, ids_r, feat_r = trk.download()
        assert n_left[0] == 0 and n_right[0] == 0                       # goodFeaturesToTrack finds nothing on a constant image, every tick
        r = ref.track_image(0.05 * (t + 1), left[t], right[t])
        n = int(n_left[1])
        assert n == len(r[0]) and np.array_equal(ids[1, :n], r[0]) and np.array_equal(cnt[1, :n], r[1])
        assert np.abs(feat[1, :n, 2:4] - r[2]).max() <= 1e-2
    # a stereo session needs a right image; the failed call changes nothing
    try:
        trk.track(0.2, np.stack([flat, left[0]]), None)
        accepted = True
    except ViwbError:
        accepted = False
    assert not accepted
    trk.track(0.2, np.stack([flat, left[0]]), np.stack([flat, right[0]]))          # the session is still usable
    assert trk.download()[0][1] > 0
    trk.close()
    # the largest point capacity the detector supports, on a mono session
    trk = ctx.tracker(1, w, h, cam, None, 1024, 2, True)
    trk.track(0.05, left[0][None])
    n_left = trk.download()[0]
    g = fo.FeatureTrackerRef(cam, None, 1024, 2, True)
    r = g.track_image(0.05, left[0])
    assert int(n_left[0]) == len(r[0]) and len(r[0]) > 150 and np.array_equal(trk.feat[0, : len(r[0]), 2:4], r[2])
    trk.close()
    for bad in (0, 1025):
        try:
            ctx.tracker(1, w, h, cam, None, bad, 10, True)
            raise AssertionError("max_cnt %d accepted" % bad)
        except ViwbError:
            pass
    # argument errors of the C entry points: reported with a message, never a crash
    import ctypes as C
    hnd = C.c_void_p()
    cfg = lib_tracker_config(30, 10, cam)
    assert ctx.lib.viwb_tracker_create(ctx.h, C.c_int(1), C.c_int(2), C.c_int(2), C.byref(cfg), C.byref(hnd)) != 0 and b"image size" in ctx.lib.viwb_last_error(ctx.h)
    assert ctx.lib.viwb_tracker_create(ctx.h, C.c_int(0), C.c_int(w), C.c_int(h), C.byref(cfg), C.byref(hnd)) != 0
    assert ctx.lib.viwb_tracker_track(None, C.c_double(0.0), None, None, C.c_int(w), None, None) != 0
    trk = ctx.tracker(1, w, h, cam, None, 30, 10, True)
    img = np.ascontiguousarray(left[0][None])
    ptr = (C.c_void_p * 1)(img.ctypes.data)
    assert ctx.lib.viwb_tracker_track(trk.hnd, C.c_double(0.05), ptr, None, C.c_int(w - 1), None, None) != 0 and b"stride" in ctx.lib.viwb_last_error(ctx.h)
    pp = np.zeros((1, 30, 2), np.float32)
    assert ctx.lib.viwb_tracker_track(trk.hnd, C.c_double(0.05), ptr, None, C.c_int(w), pp.ctypes.data_as(C.c_void_p), None) != 0
    trk.track(0.05, img); assert trk.download()[0][0] > 0                    # and the session is unharmed
    trk.close()


def lib_tracker_config(max_cnt, min_dist, cam):
    import ctypes as C
    from viwb.lib import TrackerConfig
    return TrackerConfig(max_cnt, min_dist, 1, 0, (C.c_double * 16)(*([float(v) for v in cam] * 2)))


# ------------------------------------------------------------------------------------------------ against the reference's own code
def check_marginalize_vs_reference_code(ctx, oracle, reference_code, cid):
    """The library's marginalization against MarginalizationInfo::marginalize of the reference, compiled from its sources
    (oracle/_ref): kept dimension, kept blocks and the information form, block by block (the reference's column order follows the hash
    of the parameter addresses).  The input state comes from the oracle's solve; the oracle plays no part in the comparison."""
    from test_reference_factors import _information_by_block, _shift_old
    seq = synth.Sequence(synth.make_config(cid), 1, 13)
    prob, st, _ = seq.window(0)
    a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
    for k in range(2):
        ref = reference_code.marginalize(prob, a, abi.MARGIN_OLD)
        got = ctx.marginalize(prob, a, abi.MARGIN_OLD)
        assert ref["n"] == got.n, (cid, k)
        ids0, A0, b0 = _information_by_block(ref["blocks"], ref["J"], ref["r"], _shift_old)
        ids1, A1, b1 = _information_by_block(got.blocks(), got.Jmat(), got.rvec())
        assert ids0 == ids1
        assert np.abs(A1 - A0).max() <= 1e-6 * np.abs(A0).max() and np.abs(b1 - b0).max() <= 1e-6 * np.abs(b0).max(), (cid, k)
        # the prior factor on the prior the library just produced
        x = a.copy(); x[7:10] += 0.01; x[0:3] -= 0.02
        res0, jac0 = reference_code.prior_evaluate(got, x[:abi.STATE_FIXED])
        res1, jac1 = ctx.prior_evaluate(got, x)
        assert np.abs(res0 - res1).max() <= 1e-9 * max(1.0, np.abs(res0).max()) and np.abs(jac0 - jac1).max() <= 1e-12 * max(1.0, np.abs(jac0).max())
        prob, st, _ = seq.window(k + 1, prior=got, prev_state=a)
        a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)


def time_tracker_reference(cam0, cam1, max_cnt, min_dist, seq, ticks):
    """ms per frame of trackImage() restated over cv2 (one session, one host thread) -- the CPU side of profiles/track_probe.py"""
    import os
    import sys
    import time
    import cv2
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import feature_oracle as fo
    cv2.setNumThreads(1)
    ref = fo.FeatureTrackerRef(cam0, cam1, max_cnt, min_dist, True, use_cv_detector=True)
    for t in range(2):
        ref.track_image(0.05 * (t + 1), seq[0][t], seq[1][t])
    t0 = time.perf_counter()
    for t in range(2, ticks + 2):
        ref.track_image(0.05 * (t + 1), seq[0][t], seq[1][t])
    return (time.perf_counter() - t0) / ticks * 1e3


def check_reference_estimator_on_this_backend(ctx, oracle, cid):
    """The reference's UNMODIFIED Estimator::optimization() (estimator.cpp compiled into oracle/_ref) with ceres::Solve delegated to the
    library under test (viwb_window_solve through the C ABI): vector2double, problem assembly, double2vector and the marginalization are the
    reference's own code, the solve is the CUDA backend's.  The result must be what the all-oracle Estimator::optimization() restatement
    gives: poses to the BASELINE tolerance (tight: 1e-9), the prior in information form."""
    import viw_ref
    from test_reference_factors import _as_the_estimator_holds_it, _information_by_block
    seq = synth.Sequence(synth.make_config(cid), 6, 13)
    prob, st, _ = seq.window(0)
    for k in range(2):
        st = _as_the_estimator_holds_it(st)
        a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
        ref = viw_ref.estimator_optimization_with(prob, st, lambda x: ctx.window_solve(prob, x)[0], abi.MARGIN_OLD)
        rec = ref["record"]
        assert rec["structure_mismatches"] == 0 and rec["vector2double_mismatches"] == 0 and rec["visual_row_mismatches"] == 0, (cid, k, rec)
        ep, er = synth.pose_errors(a, ref["state"])
        assert ep <= TIGHT_M and er <= TIGHT_RAD, (cid, k, ep, er)
        assert np.abs(ref["state"] - a).max() <= 1e-8 * max(1.0, np.abs(a).max())
        assert ref["n"] == q.n
        ids0, A0, b0 = _information_by_block(ref["blocks"], ref["J"], ref["r"])
        ids1, A1, b1 = _information_by_block(q.blocks(), q.Jmat(), q.rvec())
        assert ids0 == ids1 and np.abs(A1 - A0).max() <= 1e-6 * np.abs(A0).max() and np.abs(b1 - b0).max() <= 1e-6 * np.abs(b0).max(), (cid, k)
        prob, st, _ = seq.window(k + 1, prior=q, prev_state=a)


def check_reference_estimator_on_product_shim(ctx, libpath, cid, dev=False):
    """The reference's UNMODIFIED estimator.cpp compiled with <ceres/ceres.h> = the product's shim (viw-fusion_b200/host) and the reference's own
    factor objects lowered by viw-fusion_b200/host/viwb_reference_adapter.h: what a maintainer gets by changing the include path and adding one
    install call.  Against the same estimator code with the solve handed to viwb_window_solve on the ORIGINAL tables: if the shim's lowering of the
    recorded ceres::Problem reproduces the tables, the solved windows and the priors the reference's own marginalization builds from them agree."""
    import pytest
    import viw_ref
    from test_reference_factors import _as_the_estimator_holds_it, _information_by_block
    if not viw_ref.product_available(dev):
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libviw_ref_product[_dev].so")
    seq = synth.Sequence(synth.make_config(cid), 6, 13)
    prob, st, _ = seq.window(0)
    for k in range(2):
        st = _as_the_estimator_holds_it(st)
        try:
            got = viw_ref.estimator_optimization_on_product_shim(libpath, prob, st, abi.MARGIN_OLD, dev)      # dev: MarginalizationInfo::marginalize() on the device too
        except RuntimeError as ex:
            if "one library under test" in str(ex):
                pytest.skip(str(ex))
            raise
        ref = viw_ref.estimator_optimization_with(prob, st, lambda x: ctx.window_solve(prob, x)[0], abi.MARGIN_OLD)
        ep, er = synth.pose_errors(got["state"], ref["state"])
        assert ep <= TIGHT_M and er <= TIGHT_RAD, (cid, k, ep, er)
        assert np.abs(got["state"] - ref["state"]).max() <= 1e-9 * max(1.0, np.abs(ref["state"]).max()), (cid, k)
        assert got["n"] == ref["n"] and got["m"] == ref["m"]
        if dev:     # the device prior orders the kept blocks by block id, the reference by the hash of their addresses: same blocks, other columns
            assert sorted(bq for bq, _ in got["blocks"]) == sorted(bq for bq, _ in ref["blocks"])
        else:
            assert got["blocks"] == ref["blocks"]
        ids0, A0, b0 = _information_by_block(ref["blocks"], ref["J"], ref["r"])
        ids1, A1, b1 = _information_by_block(got["blocks"], got["J"], got["r"])
        tol = 1e-6 if dev else 1e-7          # (device marginalization vs the reference's CPU one: the tolerance of the marginalization parity tests)
        assert ids0 == ids1 and np.abs(A1 - A0).max() <= tol * np.abs(A0).max() and np.abs(b1 - b0).max() <= tol * np.abs(b0).max(), (cid, k)
        a, _, q = ctx.optimization(prob, st, abi.MARGIN_OLD)
        prob, st, _ = seq.window(k + 1, prior=q, prev_state=a)


def check_reference_tracker_on_this_backend(ctx, w=320, h=240, max_cnt=60, min_dist=20, ticks=5):
    """The reference's UNMODIFIED FeatureTracker::trackImage() (feature_tracker.cpp compiled into oracle/_ref) with cv::calcOpticalFlowPyrLK and
    cv::goodFeaturesToTrack answered by the library under test (viwb_lk_track, viwb_good_features_to_track) -- the replacement INTEGRATION.md
    section 2 describes -- against the same tracker code running on the real cv2: same ids in the same order, same track counts, points within
    the LK tolerance."""
    import cv2
    import viw_ref
    cam0 = (461.1586 * w / 752, 459.7529 * w / 752, w / 2 - 3.2, h / 2 + 1.7, -0.2847798, 0.08245052, -1.0946e-06, 4.78701e-06)
    cam1 = (457.5874 * w / 752, 456.1340 * w / 752, w / 2 + 4.1, h / 2 - 2.6, -0.2836831, 0.07395907, 1.9359e-04, 1.7618e-05)
    left, right, _ = camera_sequence(83, w, h, ticks)
    was = cv2.useOptimized(); cv2.setUseOptimized(False)             # goodFeaturesToTrack's scalar path is the specification (tests/test_feature_oracle.py)
    try:
        a = viw_ref.ReferenceFeatureTracker(cam0, cam1, w, h, max_cnt, min_dist, True)
        b = viw_ref.ReferenceFeatureTracker(cam0, cam1, w, h, max_cnt, min_dist, True)
        followed = 0
        for t in range(ticks):
            viw_ref.use_backend(None)
            ids0, cnt0, f0, idr0, fr0 = a.track_image(0.05 * (t + 1), left[t], right[t])
            viw_ref.use_backend(ctx)
            ids1, cnt1, f1, idr1, fr1 = b.track_image(0.05 * (t + 1), left[t], right[t])
            assert np.array_equal(ids0, ids1) and np.array_equal(cnt0, cnt1) and np.array_equal(idr0, idr1), t
            assert np.abs(f0[:, 2:4] - f1[:, 2:4]).max() <= 1e-2 and np.abs(fr0[:, 2:4] - fr1[:, 2:4]).max() <= 1e-2
            assert np.abs(f0[:, 0:2] - f1[:, 0:2]).max() <= 1.5e-2 / cam0[0]
            followed += int((cnt0 > 1).sum())
        assert followed > (ticks - 1) * max_cnt // 3
    finally:
        viw_ref.use_backend(None)
        cv2.setUseOptimized(was)

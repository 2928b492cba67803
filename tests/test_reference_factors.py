"""Pins the restated oracle to THE REFERENCE'S OWN CODE: vins_estimator/src/factor/{projection*Factor.cpp, imu_factor.h,
integration_base.h, pose_local_parameterization.cpp, pose_subset_parameterization.cpp, orientation_subset_parameterization.cpp}
compiled unmodified from /root/reference (oracle/Makefile `ref`, third-party headers replaced by oracle/refshim/) and called on
the same inputs as oracle/vo_factors.c.  Residuals and every Jacobian block must agree to round-off (measured: 3e-13 relative to the
block's scale for the projection factors, 5e-16 for the IMU factor, 1e-13 for the pre-integration records)."""
import numpy as np
import pytest

from viwb import abi, synth

vr = pytest.importorskip("viw_ref")
if not vr.available():
    pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libviw_ref.so", allow_module_level=True)

from test_oracle_factors import block_ptr, rand_pose  # noqa: E402


def close(a, b, tol=1e-11):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("cid,ftype", [(1, abi.F_PROJ_2F1C), (2, abi.F_PROJ_2F1C), (2, abi.F_PROJ_2F2C), (2, abi.F_PROJ_1F2C), (3, abi.F_PROJ_2F1C), (4, abi.F_PROJ_2F2C)])
def test_projection_factors_match_reference_code(oracle, cid, ftype):
    prob, st, _ = synth.make_window(cid)
    st = st.copy()
    st[abi.block_offset(abi.BLK_TD)] = 0.004
    idx = np.nonzero(prob.vis_type == ftype)[0][:60]
    assert len(idx) > 0
    sizes, nres = abi.FACTOR_BLOCK_SIZES[ftype], abi.FACTOR_RESIDUALS[ftype]
    for f in idx:
        fi, fj, lm = prob.vis_frame_i[f], prob.vis_frame_j[f], 32 + prob.vis_landmark[f]
        blocks = {abi.F_PROJ_2F1C: [fi, fj, 22, lm, 30], abi.F_PROJ_2F2C: [fi, fj, 22, 23, lm, 30], abi.F_PROJ_1F2C: [22, 23, lm, 30]}[ftype]
        params = [block_ptr(st, b) for b in blocks]
        r0, J0 = vr.factor_evaluate(ftype, prob.globals, prob.vis_obs[f], params, sizes, nres)
        r1, J1 = oracle.factor_evaluate(ftype, prob.globals, prob.vis_obs[f], params)
        assert close(r1, r0), (f, r0, r1)
        for a, b in zip(J1, J0):
            assert close(a, b), f
    # jacobians == NULL and jacobians[i] == NULL are honoured the same way
    r2, _ = vr.factor_evaluate(ftype, prob.globals, prob.vis_obs[idx[0]], params, sizes, nres, want_jac=False)
    r3, J3 = vr.factor_evaluate(ftype, prob.globals, prob.vis_obs[idx[0]], params, sizes, nres, null_jac=(1,))
    assert J3[1] is None and np.array_equal(r2, r3)


@pytest.mark.parametrize("cid", [1, 2, 4])
def test_imu_factor_matches_reference_code(oracle, cid):
    prob, st, _ = synth.make_window(cid)
    sizes, nres = abi.FACTOR_BLOCK_SIZES[abi.F_IMU], abi.FACTOR_RESIDUALS[abi.F_IMU]
    for f in range(len(prob.imu_frame_i)):
        i, j = prob.imu_frame_i[f], prob.imu_frame_j[f]
        params = [block_ptr(st, b) for b in [i, 11 + i, j, 11 + j]]
        r0, J0 = vr.factor_evaluate(abi.F_IMU, prob.globals, prob.imu_data[f], params, sizes, nres)
        r1, J1 = oracle.factor_evaluate(abi.F_IMU, prob.globals, prob.imu_data[f], params)
        assert close(r1, r0, 1e-12), (f, np.abs(r1 - r0).max())
        for a, b in zip(J1, J0):
            assert close(a, b, 1e-12), f


def test_preintegration_matches_reference_code(oracle):
    rng = np.random.default_rng(12)
    noise = np.array([0.1, 0.01, 1e-3, 1e-4])
    for n in (1, 7, 10, 40):
        dt = np.full(n, 0.005) * rng.uniform(0.8, 1.2, n)
        acc = rng.normal(0, 1.0, (n + 1, 3)) + [0, 0, 9.8]
        gyr = rng.normal(0, 0.4, (n + 1, 3))
        ba, bg = rng.normal(0, 0.05, 3), rng.normal(0, 0.01, 3)
        r0 = vr.imu_preintegrate(dt, acc, gyr, ba, bg, noise)
        r1 = oracle.imu_preintegrate(dt, acc, gyr, ba, bg, noise)
        assert np.abs(r1[:62] - r0[:62]).max() <= 1e-12 * max(1.0, np.abs(r0[:62]).max()), n
        assert np.abs(r1[62:] - r0[62:]).max() <= 1e-12 * np.abs(r0[62:]).max(), n
        r2 = synth.imu_preintegrate(dt, acc, gyr, ba, bg, noise)                 # the generator's own restatement, too
        assert np.abs(r2 - r0).max() <= 1e-10 * max(1.0, np.abs(r0).max())


def test_manifolds_match_reference_code(oracle):
    rng = np.random.default_rng(13)
    prob, st, _ = synth.make_window(4)
    for trial in range(20):
        x = rand_pose(rng)
        d = rng.normal(0, 0.05, 6)
        for kind, mask in [(0, 0), (1, 1 << 2), (1, (1 << 2) | (1 << 5)), (1, 0b111000)]:
            out, jac = vr.manifold(kind, mask, x, d)
            # the oracle's Plus on a full state: put x into pose block 0, apply delta on its tangent columns
            s2 = st.copy(); s2[0:7] = x
            delta = np.zeros(abi.TANGENT_FIXED + prob.num_landmarks); delta[0:6] = d
            prob.subset_mask[0] = mask
            got = oracle.state_plus(prob, s2, delta)[0:7]
            prob.subset_mask[0] = 0
            assert np.abs(got - out).max() <= 1e-14, (kind, mask)
            assert np.array_equal(jac, np.vstack([np.eye(6), np.zeros((1, 6))]))       # ComputeJacobian = [I; 0] whatever the mask (quirk 2)
        q = x[3:7]
        out, jac = vr.manifold(2, 1 << 2, q, d[:3])
        assert abs(np.linalg.norm(out) - 1.0) < 1e-12 and np.array_equal(jac, np.vstack([np.eye(3), np.zeros((1, 3))]))


@pytest.mark.parametrize("dtd,equal_gyr", [(0.0, False), (0.004, False), (-0.003, True)])
def test_wheel_factor_matches_reference_code(oracle, dtd, equal_gyr):
    """WheelFactor::Evaluate (factor/wheel_factor.h:28-246) with WheelIntegrationBase::evaluate (wheel_integration_base.h:179-218)
    and the reference's own sophus_utils.hpp, all compiled from the reference tree; quirk 4 (inexact sx/sy/sw/td Jacobians for
    dtd != 0) must come out the same on both sides."""
    rng = np.random.default_rng(21)
    sizes, nres = abi.FACTOR_BLOCK_SIZES[abi.F_WHEEL], abi.FACTOR_RESIDUALS[abi.F_WHEEL]
    for cid in (3, 4):
        prob, st, _ = synth.make_window(cid)
        for f in range(len(prob.wheel_frame_i)):
            i, j = prob.wheel_frame_i[f], prob.wheel_frame_j[f]
            params = [block_ptr(st, b) for b in [i, j, 24, 27, 28, 29, 31]]
            params[3][0], params[4][0], params[5][0] = 1.0 + rng.normal(0, 0.01), 1.0 + rng.normal(0, 0.01), 1.0 + rng.normal(0, 0.01)
            params[6][0] = dtd
            rec = prob.wheel_data[f].copy()
            if equal_gyr:
                rec[74:77] = rec[68:71]
            r0, J0 = vr.factor_evaluate(abi.F_WHEEL, prob.globals, rec, params, sizes, nres)
            r1, J1 = oracle.factor_evaluate(abi.F_WHEEL, prob.globals, rec, params)
            assert close(r1, r0, 1e-10), (cid, f, np.abs(r1 - r0).max())
            for bi, (a, b) in enumerate(zip(J1, J0)):
                assert close(a, b, 1e-10), (cid, f, bi, np.abs(a - b).max())


def test_plane_factor_matches_reference_code(oracle):
    rng = np.random.default_rng(22)
    from viwb.geom import R_to_q, so3_exp
    prob, st, _ = synth.make_window(4)
    sizes, nres = abi.FACTOR_BLOCK_SIZES[abi.F_PLANE], abi.FACTOR_RESIDUALS[abi.F_PLANE]
    for trial in range(20):
        params = [rand_pose(rng), rand_pose(rng, 0.3), R_to_q(so3_exp(rng.normal(0, 0.2, 3))), np.array([rng.normal()])]
        r0, J0 = vr.factor_evaluate(abi.F_PLANE, prob.globals, None, params, sizes, nres)
        r1, J1 = oracle.factor_evaluate(abi.F_PLANE, prob.globals, None, params)
        assert close(r1, r0, 1e-11)
        for a, b in zip(J1, J0):
            assert close(a, b, 1e-11), trial
    # and on the window's own plane factors
    for f in prob.plane_frame[:5]:
        params = [block_ptr(st, b) for b in [int(f), 24, 25, 26]]
        r0, J0 = vr.factor_evaluate(abi.F_PLANE, prob.globals, None, params, sizes, nres)
        r1, J1 = oracle.factor_evaluate(abi.F_PLANE, prob.globals, None, params)
        assert close(r1, r0, 1e-11) and all(close(a, b, 1e-11) for a, b in zip(J1, J0))


def test_wheel_preintegration_matches_reference_code(oracle):
    rng = np.random.default_rng(23)
    noise = np.array([0.01, 0.004])
    for n in (1, 3, 10):
        dt = np.full(n, 0.02) * rng.uniform(0.8, 1.2, n)
        vel = rng.normal(0, 0.05, (n + 1, 3)) + [0.8, 0.0, 0.0]
        gyr = rng.normal(0, 0.2, (n + 1, 3))
        s, td = 1.0 + rng.normal(0, 0.02, 3), 0.002
        r0 = vr.wheel_preintegrate(dt, vel, gyr, s, td, noise)
        r1 = oracle.wheel_preintegrate(dt, vel, gyr, s, td, noise)
        assert np.abs(r1 - r0).max() <= 1e-12 * max(1.0, np.abs(r0).max()), (n, np.abs(r1 - r0).max())
        r2 = synth.wheel_preintegrate(dt, vel, gyr, s, td, noise)
        assert np.abs(r2 - r0).max() <= 1e-10 * max(1.0, np.abs(r0).max())


# ------------------------------------------------------------------------------------------------ prior factor + marginalization
def _information_by_block(blocks, J, r, remap=None):
    """(A, b) = (J^T J, J^T r) re-ordered into ascending block id, so that two priors whose columns are ordered differently
    (the reference orders them by the hash of the parameter addresses) can be compared entry by entry."""
    A, b = J.T @ J, J.T @ r
    cols, ids = [], []
    for bid, idx in sorted(((remap(b_) if remap else b_), i_) for b_, i_ in blocks):
        size = abi.block_tsize(bid) if abi.block_size(bid) != 7 else 6
        cols += list(range(idx, idx + size)); ids.append(bid)
    return ids, A[np.ix_(cols, cols)], b[cols]


def _shift_old(b):          # addr_shift of MARGIN_OLD (estimator.cpp:1791-1800): pose i -> i-1, speed-bias i -> i-1
    return b - 1 if (1 <= b <= 10 or 12 <= b <= 21) else b


@pytest.mark.parametrize("cid", [1, 2, 3, 4, 6])
def test_marginalization_matches_reference_code(oracle, cid):
    """MARGIN_OLD on a first window (no prior) and on the next one (prior present): the reference's ResidualBlockInfo::Evaluate (loss
    corrector), preMarginalize and marginalize (4 threads, Schur complement, eigen-decomposition with the 1e-8 cut) against the oracle."""
    seq = synth.Sequence(synth.make_config(cid), 1, 13)
    prob, st, _ = seq.window(0)
    a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
    for k in range(2):
        ref = vr.marginalize(prob, a, abi.MARGIN_OLD)
        got = oracle.marginalize(prob, a, abi.MARGIN_OLD)
        assert ref["n"] == got.n, (cid, k, ref["n"], got.n)
        ids0, A0, b0 = _information_by_block(ref["blocks"], ref["J"], ref["r"], _shift_old)
        ids1, A1, b1 = _information_by_block(got.blocks(), got.Jmat(), got.rvec())
        assert ids0 == ids1, (ids0, ids1)
        sA, sb = np.abs(A0).max(), np.abs(b0).max()
        assert np.abs(A1 - A0).max() <= 1e-7 * sA and np.abs(b1 - b0).max() <= 1e-7 * sb, (cid, k, np.abs(A1 - A0).max() / sA, np.abs(b1 - b0).max() / sb)
        # next window: the prior just produced takes part, with pose 0 / speed-bias 0 dropped from it
        prob, st, _ = seq.window(k + 1, prior=got, prev_state=a)
        a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)


@pytest.mark.parametrize("cid", [2, 4])
def test_prior_factor_matches_reference_code(oracle, cid):
    seq = synth.Sequence(synth.make_config(cid), 2, 13)
    prob, st, _ = seq.window(0)
    a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
    prob1, st1, _ = seq.window(1, prior=q, prev_state=a)
    rng = np.random.default_rng(31)
    for trial in range(4):
        x = st1.copy()
        x[:abi.STATE_FIXED] += rng.normal(0, 0.01 * trial, abi.STATE_FIXED)
        for f in range(11):
            x[7 * f + 3: 7 * f + 7] /= np.linalg.norm(x[7 * f + 3: 7 * f + 7])
        if trial == 3:                                   # the w < 0 branch of the quaternion difference (marginalization_factor.cpp:374-377)
            x[3:7] *= -1.0
        r0, J0 = vr.prior_evaluate(q, x)
        r1, J1 = oracle.prior_evaluate(q, x)
        assert close(r1, r0, 1e-12) and close(J1, J0, 1e-12), trial


def _shift_second_new(b):   # addr_shift of MARGIN_SECOND_NEW (estimator.cpp:1843-1870): frame 10 takes the place of frame 9
    return 9 if b == 10 else 20 if b == 21 else b


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_marginalization_second_new_matches_reference_code(oracle, cid):
    seq = synth.Sequence(synth.make_config(cid), 3, 13)
    prob, st, _ = seq.window(0)
    a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
    prob, st, _ = seq.window(1, prior=q, prev_state=a)
    a, sm, _ = oracle.optimization(prob, st, abi.MARGIN_SECOND_NEW, want_prior=False)
    ref = vr.marginalize(prob, a, abi.MARGIN_SECOND_NEW)
    got = oracle.marginalize(prob, a, abi.MARGIN_SECOND_NEW)
    assert ref["n"] == got.n
    ids0, A0, b0 = _information_by_block(ref["blocks"], ref["J"], ref["r"], _shift_second_new)
    ids1, A1, b1 = _information_by_block(got.blocks(), got.Jmat(), got.rvec())
    assert ids0 == ids1
    assert np.abs(A1 - A0).max() <= 1e-7 * np.abs(A0).max() and np.abs(b1 - b0).max() <= 1e-7 * np.abs(b0).max()


# ------------------------------------------------------------------------------------------------ FeatureManager
def test_triangulation_and_depth_shift_match_reference_code(oracle):
    """FeatureManager::triangulate (stereo and two-frame branches, triangulatePoint's 4x4 SVD) and removeBackShiftDepth from
    estimator/feature_manager.cpp, compiled unmodified (cv::solvePnP stubbed out, not exercised), against the oracle's restatement."""
    from viwb import geom
    rng = np.random.default_rng(51)
    prob, st, gt = synth.Sequence(synth.make_config(2), 1, 11).window(0)
    x = gt.copy()
    n = 80
    stereo = (rng.uniform(size=n) < 0.5).astype(np.int32)
    frame = rng.integers(0, 9, n).astype(np.int32)

    def cam(i, c):
        Rs = geom.q_to_R(x[7 * i + 3: 7 * i + 7]); ric = geom.q_to_R(x[176 + 7 * c + 3: 176 + 7 * c + 7])
        R = Rs @ ric; t = x[7 * i: 7 * i + 3] + Rs @ x[176 + 7 * c: 176 + 7 * c + 3]
        return np.hstack([R.T, (-R.T @ t)[:, None]])
    pt0, pt1 = np.zeros((n, 2)), np.zeros((n, 2))
    for k in range(n):
        P0 = cam(frame[k], 0); P1 = cam(frame[k], 1) if stereo[k] else cam(frame[k] + 1, 0)
        pc = np.array([rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), 1.0]) * rng.uniform(2, 15)
        if k % 10 == 9:
            pc[2] = -abs(pc[2])                                           # behind the camera: the INIT_DEPTH branch
        pw = P0[:, :3].T @ (pc - P0[:, 3])
        q1 = P1[:, :3] @ pw + P1[:, 3]
        pt0[k] = pc[:2] / pc[2] + rng.normal(0, 1e-3, 2); pt1[k] = q1[:2] / q1[2] + rng.normal(0, 1e-3, 2)
    ref = vr.triangulate(x, stereo, frame, pt0, pt1)
    got = oracle.triangulate(x, stereo, frame, pt0, pt1)
    assert (ref == 5.0).sum() >= 4 and np.array_equal(ref == 5.0, got == 5.0)
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max(), np.abs(got - ref).max()
    uv = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-0.6, 0.6, n), np.ones(n)])
    dep = rng.uniform(1, 20, n); dep[:3] = -1.0
    mR, nR = geom.q_to_R(x[3:7]) @ geom.q_to_R(x[179:183]), geom.q_to_R(x[10:14]) @ geom.q_to_R(x[179:183])
    mP, nP = x[0:3] + geom.q_to_R(x[3:7]) @ x[176:179], x[7:10] + geom.q_to_R(x[10:14]) @ x[176:179]
    r0 = vr.shift_depth(uv, dep, mR, mP, nR, nP)
    r1 = oracle.shift_depth(uv, dep, mR, mP, nR, nP)
    assert np.abs(r1 - r0).max() <= 1e-13 * np.abs(r0).max()


# ------------------------------------------------------------------------------------------------ camera model
def test_undistorted_pts_match_reference_code():
    """FeatureTracker::undistortedPts = PinholeCamera::liftProjective (recursive distortion model, 8 iterations) from
    camera_models/src/camera_models/PinholeCamera.cc, compiled unmodified, against the checker's restatement: bit for bit."""
    import feature_oracle as fo
    rng = np.random.default_rng(61)
    for cam, (w, h) in (((461.1586, 459.7529, 362.6593, 248.5236, -0.2847798, 0.08245052, -1.0946e-06, 4.78701e-06), (752, 480)),
                        ((457.5874, 456.1340, 379.9994, 255.2381, -0.2836831, 0.07395907, 1.9359e-04, 1.7618e-05), (752, 480)),
                        ((384.45, 384.45, 320.0, 240.0, 0.0, 0.0, 0.0, 0.0), (640, 480))):                   # the m_noDistortion branch
        pts = np.column_stack([rng.uniform(0, w, 400), rng.uniform(0, h, 400)]).astype(np.float32)
        assert np.array_equal(vr.undistorted_pts(cam, w, h, pts), fo.lift_projective(cam, pts))


# ------------------------------------------------------------------------------------------------ Estimator::optimization() itself
def _as_the_estimator_holds_it(st):
    """The reference keeps rotations as matrices and vector2double() turns them into quaternions (estimator.cpp:1155-1222, Eigen's
    matrix -> quaternion: the trace branch, else the largest diagonal entry), so of the two quaternions of a rotation the arrays always hold the
    one that conversion yields; the synthetic generator's perturbed quaternions may be the other.  Same rotation, the representative the
    reference would hand to the solver."""
    from viwb.geom import q_to_R
    st = st.copy()
    for off in [7 * i + 3 for i in range(11)] + [179, 186, 193, 197]:
        q = st[off: off + 4] / np.linalg.norm(st[off: off + 4])
        m = q_to_R(q)
        t = m[0, 0] + m[1, 1] + m[2, 2]
        out = np.zeros(4)
        if t > 0:
            t = np.sqrt(t + 1.0); out[3] = 0.5 * t; t = 0.5 / t
            out[0], out[1], out[2] = (m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t
        else:
            a = 0
            if m[1, 1] > m[0, 0]:
                a = 1
            if m[2, 2] > m[a, a]:
                a = 2
            b, c = (a + 1) % 3, (a + 2) % 3
            t = np.sqrt(m[a, a] - m[b, b] - m[c, c] + 1.0); out[a] = 0.5 * t; t = 0.5 / t
            out[3] = (m[c, b] - m[b, c]) * t; out[b] = (m[b, a] + m[a, b]) * t; out[c] = (m[c, a] + m[a, c]) * t
        if np.dot(out, q) < 0:                      # keep the handed-in digits, only pick the representative
            st[off: off + 4] *= -1.0
    return st


@pytest.mark.parametrize("cid", [1, 2, 3, 4, 6])
def test_estimator_optimization_of_the_reference_runs_on_the_tables(oracle, cid):
    """estimator/estimator.cpp compiled unmodified (ROS / OpenCV / camodocal / ceres as name stand-ins, oracle/refshim/): the reference's own
    Estimator::optimization() is driven on two consecutive windows.  ceres::Problem records, ceres::Solve checks the record against the
    tables and plays the restated solver's solution back (Ceres itself is not in the image).  Pinned thereby:
      a-1  vector2double -- the arrays the solver sees are the window handed in; double2vector -- the re-anchored window equals the oracle's;
      a-2  the problem assembly -- every parameter block (presence, constancy, manifold and its subset mask) and every residual block (type,
           blocks, constants, order) equals the table row the library consumes;
      a-13 the marginalization as optimization() orchestrates it (factor hand-over, drop sets, address shift) -- the new prior."""
    seq = synth.Sequence(synth.make_config(cid), 4, 13)
    prob, st, _ = seq.window(0)
    for k in range(2):
        st = _as_the_estimator_holds_it(st)
        solved, sm = oracle.window_solve(prob, st)
        a, sm2, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
        ref = vr.estimator_optimization(prob, st, solved, abi.MARGIN_OLD)
        rec = ref["record"]
        assert rec["structure_mismatches"] == 0 and rec["vector2double_mismatches"] == 0 and rec["visual_row_mismatches"] == 0, (cid, k, rec)
        assert rec["prior"] == (1 if prob.prior is not None and prob.prior.valid else 0)
        assert rec["imu"] == len(prob.imu_frame_i) and rec["wheel"] == len(prob.wheel_frame_i) and rec["plane"] == len(prob.plane_frame)
        assert [rec["proj_2f1c"], rec["proj_2f2c"], rec["proj_1f2c"]] == [int((prob.vis_type == t).sum()) for t in (0, 1, 2)]
        assert np.abs(ref["state"] - a).max() <= 1e-12 * max(1.0, np.abs(a).max()), (cid, k, np.abs(ref["state"] - a).max())
        assert ref["n"] == q.n
        ids0, A0, b0 = _information_by_block(ref["blocks"], ref["J"], ref["r"])
        ids1, A1, b1 = _information_by_block(q.blocks(), q.Jmat(), q.rvec())
        assert ids0 == ids1
        assert np.abs(A1 - A0).max() <= 1e-7 * np.abs(A0).max() and np.abs(b1 - b0).max() <= 1e-7 * np.abs(b0).max(), (cid, k)
        prob, st, _ = seq.window(k + 1, prior=q, prev_state=a)


@pytest.mark.parametrize("cid", [2, 4, 6])
def test_estimator_optimization_second_new(oracle, cid):
    seq = synth.Sequence(synth.make_config(cid), 5, 13)
    prob, st, _ = seq.window(0)
    a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
    prob, st, _ = seq.window(1, prior=q, prev_state=a)
    st = _as_the_estimator_holds_it(st)
    solved, sm = oracle.window_solve(prob, st)
    a, sm2, q2 = oracle.optimization(prob, st, abi.MARGIN_SECOND_NEW)
    ref = vr.estimator_optimization(prob, st, solved, abi.MARGIN_SECOND_NEW)
    assert ref["record"]["structure_mismatches"] == 0 and ref["record"]["visual_row_mismatches"] == 0
    assert np.abs(ref["state"] - a).max() <= 1e-12 * max(1.0, np.abs(a).max())
    assert ref["n"] == q2.n
    ids0, A0, b0 = _information_by_block(ref["blocks"], ref["J"], ref["r"])
    ids1, A1, b1 = _information_by_block(q2.blocks(), q2.Jmat(), q2.rvec())
    assert ids0 == ids1 and np.abs(A1 - A0).max() <= 1e-7 * np.abs(A0).max() and np.abs(b1 - b0).max() <= 1e-7 * np.abs(b0).max()


@pytest.mark.parametrize("cid", [1, 2, 4, 6])
def test_outlier_rejection_matches_reference_code(oracle, cid):
    """Estimator::outliersRejection / reprojectionError (estimator.cpp:2115-2185) of the reference on solved and on corrupted windows"""
    prob, st, _ = synth.make_window(cid)
    a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
    rng = np.random.default_rng(71)
    for trial in range(3):
        x = a.copy()
        if trial:
            bad = rng.choice(prob.num_landmarks, 12, replace=False)
            x[abi.STATE_FIXED + bad] *= rng.uniform(0.2, 3.0, 12)              # wrong depths: large reprojection errors
        r0 = vr.estimator_outliers(prob, x)
        r1 = oracle.outlier_rejection(prob, x)
        assert np.array_equal(r0, r1), (cid, trial, int(r0.sum()), int(r1.sum()))
        assert trial == 0 or r0.sum() >= 3


# ------------------------------------------------------------------------------------------------ FeatureTracker::trackImage
def test_track_image_restatement_matches_reference_code():
    """featureTracker/feature_tracker.cpp compiled unmodified, its three OpenCV calls (calcOpticalFlowPyrLK, goodFeaturesToTrack, circle)
    answered by the real cv2 through callbacks: the reference's own trackImage() against the checker's line-by-line restatement
    (oracle/feature_oracle.py:FeatureTrackerRef, the one the device tracker is tested against).  setMask() sorts with std::sort, which
    leaves the order of equal track counts unspecified; for this test the restatement visits them in the order this libstdc++ produces
    (vr.std_sort_order; the device tracker documents a stable order, an equally valid choice) and everything must then be IDENTICAL, row by
    row: ids, track counts, pixels, undistorted points, velocities, both cameras, every tick."""
    import feature_oracle as fo
    from parity_checks import camera_sequence
    cv2 = pytest.importorskip("cv2")
    w, h, max_cnt, min_dist, ticks = 320, 240, 60, 20, 6
    cam0 = (461.1586 * w / 752, 459.7529 * w / 752, w / 2 - 3.2, h / 2 + 1.7, -0.2847798, 0.08245052, -1.0946e-06, 4.78701e-06)
    cam1 = (457.5874 * w / 752, 456.1340 * w / 752, w / 2 + 4.1, h / 2 - 2.6, -0.2836831, 0.07395907, 1.9359e-04, 1.7618e-05)
    was = cv2.useOptimized(); cv2.setUseOptimized(False)
    try:
        for stereo in (True, False):
            left, right, motion = camera_sequence(81, w, h, ticks)
            ref = vr.ReferenceFeatureTracker(cam0, cam1 if stereo else None, w, h, max_cnt, min_dist, True)
            mine = fo.FeatureTrackerRef(cam0, cam1 if stereo else None, max_cnt, min_dist, True, use_cv_detector=True, order_fn=vr.std_sort_order)
            followed = 0
            for t in range(ticks):
                if t in (2, 4):                                          # the hasPrediction branch: a good prediction, then one outside the image (< 10 successes -> full-pyramid repeat)
                    pred = (mine.prev_pts + (motion[t] if t == 2 else np.float32([w + 30.0, h + 30.0]))).astype(np.float32)
                    ref.set_prediction(pred); mine.set_prediction(pred)
                ids0, cnt0, f0, idr0, fr0 = ref.track_image(0.05 * (t + 1), left[t], right[t] if stereo else None)
                ids1, cnt1, pts1, un1, vel1, idr1, ptsr1, unr1, velr1 = mine.track_image(0.05 * (t + 1), left[t], right[t] if stereo else None)
                assert np.array_equal(ids0, ids1), (stereo, t)
                o0, o1 = np.argsort(ids0), np.argsort(ids1)
                assert np.array_equal(cnt0[o0], cnt1[o1])
                assert np.array_equal(f0[o0, 2:4], pts1[o1]) and np.array_equal(f0[o0, 0:2], un1[o1]) and np.array_equal(f0[o0, 4:6], vel1[o1]), (stereo, t)
                assert sorted(idr0) == sorted(idr1)
                if stereo:
                    p0, p1 = np.argsort(idr0), np.argsort(idr1)
                    assert np.array_equal(fr0[p0, 2:4], ptsr1[p1]) and np.array_equal(fr0[p0, 0:2], unr1[p1]) and np.array_equal(fr0[p0, 4:6], velr1[p1])
                followed += int((cnt0 > 1).sum())
            assert followed > (ticks - 1) * max_cnt // 3
            assert mine.stats == {"predicted": 2, "repeated": 1}
    finally:
        cv2.setUseOptimized(was)


@pytest.mark.parametrize("cid", [1, 2, 3, 4])
def test_visual_imu_alignment_restatement_matches_reference_code(oracle, cid):
    """SURVEY 8 f-4 ii: VisualIMUAlignment of the reference (initial/initial_aligment.cpp compiled unmodified: solveGyroscopeBias with its
    repropagation, LinearAlignment[WithWheel], RefineGravity[WithWheel]) against the numpy restatement oracle/init_oracle.py -- gyroscope
    bias, the repropagated pre-integration records, gravity, the solution vector and the verdict, mono / stereo / wheel shapes."""
    import init_oracle as io
    import parity_checks as pc
    case = pc.alignment_case(cid, F=13, seed=5)
    n = len(case["dts"])
    wheel = None
    if case["cfg"].use_wheel:
        w = case["wheel_samples"]
        wheel = np.array([oracle.wheel_preintegrate(i[0], i[1], i[2], np.ones(3), 0.0, np.array([case["cfg"].vel_n_wheel, case["cfg"].gyr_n_wheel])) for i in w])
    ref = vr.visual_imu_alignment(case["R"], case["T"], case["dts"], case["accs"], case["gyrs"], case["noise"], case["bg0"], wheel, case["tic"], case["rio"], case["tio"],
                                  np.array([0, 0, case["g_norm"]]))
    rec0 = np.array([oracle.imu_preintegrate(case["dts"][i], case["accs"][i], case["gyrs"][i], np.zeros(3), case["bg0"], case["noise"]) for i in range(n)])
    dbg = io.solve_gyroscope_bias(case["R"], rec0)
    assert close(dbg, ref["delta_bg"], 1e-10)
    rec = np.array([oracle.imu_preintegrate(case["dts"][i], case["accs"][i], case["gyrs"][i], np.zeros(3), case["bg0"] + dbg, case["noise"]) for i in range(n)])
    assert close(rec, ref["imu"], 1e-10)
    ok, g, x = io.linear_alignment(case["R"], case["T"], rec, wheel, case["tic"], case["rio"], case["tio"], case["g_norm"])
    assert ok == ref["ok"] and ok and len(x) == len(ref["x"])
    assert close(g, ref["g"], 1e-8) and close(x, ref["x"], 1e-8), (np.abs(g - ref["g"]).max(), np.abs(x - ref["x"]).max())
    # the rejected branch: |g| off by more than 0.5 -> false, the unrefined 3F+4 vector
    bad = vr.visual_imu_alignment(case["R"], case["T"], case["dts"], case["accs"], case["gyrs"], case["noise"], case["bg0"], wheel, case["tic"], case["rio"], case["tio"],
                                  np.array([0, 0, case["g_norm"] + 2.0]))
    okb, gb, xb = io.linear_alignment(case["R"], case["T"], rec, wheel, case["tic"], case["rio"], case["tio"], case["g_norm"] + 2.0)
    assert bad["ok"] == okb and not okb and len(bad["x"]) == len(xb) == 3 * len(case["R"]) + 4 and close(xb, bad["x"], 1e-8)

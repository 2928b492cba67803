"""Pins the CPU oracle's factor restatements with the reference's own sanctioned check: analytic Jacobian vs
forward difference with eps = 1e-6 on the manifold (x * deltaQ(d)), projectionTwoFrameTwoCamFactor.cpp:237-303.
Known, intentional analytic != numeric entries are whitelisted."""
import numpy as np
import pytest

from viwb import abi, synth
from viwb.geom import q_mul, q_normalize, R_to_q, so3_exp

EPS = 1e-6


def plus(block, size, d):
    """x (+) d with the reference's first-order deltaQ (utility.h:22-36)."""
    x = np.array(block, float)
    if size == 7:
        x[:3] += d[:3]
        dq = q_normalize(np.r_[d[3:6] / 2, 1.0])
        x[3:7] = q_mul(x[3:7], dq)
    elif size == 4:
        dq = q_normalize(np.r_[d[:3] / 2, 1.0])
        x[:4] = q_mul(x[:4], dq)
    else:
        x[:size] += d[:size]
    return x


def numeric_jacobian(oracle, ftype, g, consts, params):
    sizes = abi.FACTOR_BLOCK_SIZES[ftype]
    r0, _ = oracle.factor_evaluate(ftype, g, consts, params, want_jac=False)
    out = []
    for i, s in enumerate(sizes):
        ts = 6 if s == 7 else 3 if s == 4 else s
        J = np.zeros((len(r0), ts))
        for k in range(ts):
            d = np.zeros(ts)
            d[k] = EPS
            p2 = [np.array(p, float) for p in params]
            p2[i] = plus(params[i], s, d)
            r1, _ = oracle.factor_evaluate(ftype, g, consts, p2, want_jac=False)
            J[:, k] = (r1 - r0) / EPS
        out.append(J)
    return out


def rand_pose(rng, scale=1.0):
    return np.r_[rng.normal(0, scale, 3), R_to_q(so3_exp(rng.normal(0, 0.3, 3)))]


def window_factor_cases(cid):
    prob, st, gt = synth.make_window(cid)
    return prob, st


def block_ptr(st, b):
    o = abi.block_offset(b) if b < 32 else abi.STATE_FIXED + b - 32
    s = abi.block_size(b) if b < 32 else 1
    return st[o:o + s].copy()


@pytest.mark.parametrize("cid,ftype", [(1, abi.F_PROJ_2F1C), (2, abi.F_PROJ_2F2C), (2, abi.F_PROJ_1F2C)])
def test_projection_jacobians(oracle, cid, ftype):
    prob, st = window_factor_cases(cid)
    st = st.copy()
    st[abi.block_offset(abi.BLK_TD)] = 0.003   # exercise the td terms
    idx = np.nonzero(prob.vis_type == ftype)[0][:25]
    assert len(idx) > 0
    for f in idx:
        fi, fj, lm = prob.vis_frame_i[f], prob.vis_frame_j[f], 32 + prob.vis_landmark[f]
        blocks = {abi.F_PROJ_2F1C: [fi, fj, 22, lm, 30], abi.F_PROJ_2F2C: [fi, fj, 22, 23, lm, 30], abi.F_PROJ_1F2C: [22, 23, lm, 30]}[ftype]
        params = [block_ptr(st, b) for b in blocks]
        r, J = oracle.factor_evaluate(ftype, prob.globals, prob.vis_obs[f], params)
        Jn = numeric_jacobian(oracle, ftype, prob.globals, prob.vis_obs[f], params)
        for i, (Ja, Jnum) in enumerate(zip(J, Jn)):
            ts = Jnum.shape[1]
            if ftype == abi.F_PROJ_1F2C and i == 2:
                continue  # quirk 3: d r / d lambda uses pts_i instead of pts_i_td (projectionOneFrameTwoCamFactor.cpp:119)
            scale = max(1.0, np.abs(Jnum).max())
            assert np.abs(Ja[:, :ts] - Jnum).max() < 2e-4 * scale, (f, i)
            if Ja.shape[1] == 7:
                assert np.all(Ja[:, 6] == 0)


def test_projection_1f2c_lambda_quirk(oracle):
    """quirk 3 is reproduced: with td == td_i the inverse-depth Jacobian is exact, otherwise it uses pts_i."""
    prob, st = window_factor_cases(2)
    f = int(np.nonzero(prob.vis_type == abi.F_PROJ_1F2C)[0][0])
    params = [block_ptr(st, b) for b in [22, 23, 32 + prob.vis_landmark[f], 30]]
    r, J = oracle.factor_evaluate(abi.F_PROJ_1F2C, prob.globals, prob.vis_obs[f], params)
    Jn = numeric_jacobian(oracle, abi.F_PROJ_1F2C, prob.globals, prob.vis_obs[f], params)
    assert np.abs(J[2] - Jn[2]).max() < 2e-4 * max(1.0, np.abs(Jn[2]).max())


def test_imu_jacobians(oracle):
    prob, st = window_factor_cases(1)
    for f in range(len(prob.imu_frame_i)):
        i, j = prob.imu_frame_i[f], prob.imu_frame_j[f]
        params = [block_ptr(st, b) for b in [i, 11 + i, j, 11 + j]]
        r, J = oracle.factor_evaluate(abi.F_IMU, prob.globals, prob.imu_data[f], params)
        Jn = numeric_jacobian(oracle, abi.F_IMU, prob.globals, prob.imu_data[f], params)
        for bi, (Ja, Jnum) in enumerate(zip(J, Jn)):
            ts = Jnum.shape[1]
            # whitening makes entries O(1e5); compare relative to the row scale.
            # (R,BG) uses the un-corrected delta_q (quirk 5) -> first-order accurate only.
            scale = np.maximum(1.0, np.abs(Jnum).max(axis=1, keepdims=True))
            assert (np.abs(Ja[:, :ts] - Jnum) / scale).max() < 5e-3, (f, bi)


def test_imu_residual_zero_at_consistent_state(oracle):
    """A state propagated with the pre-integrated deltas gives a (near) zero un-whitened residual."""
    prob, st, gt = synth.make_window(1, perturb=False)
    rec = prob.imu_data[0]
    params = [block_ptr(gt, b) for b in [0, 11, 1, 12]]
    r, _ = oracle.factor_evaluate(abi.F_IMU, prob.globals, rec, params, want_jac=False)
    # whitened residual should be O(few sigma)
    assert np.all(np.isfinite(r)) and np.abs(r).max() < 60.0


@pytest.mark.parametrize("dtd,equal_gyr", [(0.0, False), (0.01, True), (0.01, False)])
def test_wheel_jacobians(oracle, dtd, equal_gyr):
    """dtd = td_wheel - linearized_td.  The reference's d r_theta / d td uses Jr(-sw*w0*dtd) where the exact
    expression needs Jr(-sw*w1*dtd) (wheel_factor.h:188-189,236): exact iff the first and last gyro samples
    agree, so the td block is checked with gyr_1 := linearized_gyr and whitelisted otherwise.  sx/sy/sw are
    inexact for dtd != 0 (quirk 4).  Every shipped config has estimate_td_wheel = 0, i.e. dtd == 0."""
    prob, st = window_factor_cases(4)
    rng = np.random.default_rng(7)
    for f in range(3):
        i, j = prob.wheel_frame_i[f], prob.wheel_frame_j[f]
        params = [block_ptr(st, b) for b in [i, j, 24, 27, 28, 29, 31]]
        params[3][0], params[4][0], params[5][0] = 1.0 + rng.normal(0, 0.01), 1.0 + rng.normal(0, 0.01), 1.0 + rng.normal(0, 0.01)
        params[6][0] = dtd
        rec = prob.wheel_data[f].copy()
        if equal_gyr:
            rec[74:77] = rec[68:71]
        r, J = oracle.factor_evaluate(abi.F_WHEEL, prob.globals, rec, params)
        Jn = numeric_jacobian(oracle, abi.F_WHEEL, prob.globals, rec, params)
        for bi, (Ja, Jnum) in enumerate(zip(J, Jn)):
            ts = Jnum.shape[1]
            if dtd != 0.0 and bi in (3, 4, 5):
                continue
            if dtd != 0.0 and bi == 6 and not equal_gyr:
                continue
            scale = np.maximum(1.0, np.abs(Jnum).max(axis=1, keepdims=True))
            assert (np.abs(Ja[:, :ts] - Jnum) / scale).max() < 1e-4, (f, bi, dtd)


def test_plane_jacobians(oracle):
    prob, st = window_factor_cases(4)
    rng = np.random.default_rng(3)
    for f in range(3):
        params = [rand_pose(rng), rand_pose(rng, 0.3), R_to_q(so3_exp(rng.normal(0, 0.2, 3))), np.array([rng.normal()])]
        r, J = oracle.factor_evaluate(abi.F_PLANE, prob.globals, None, params)
        Jn = numeric_jacobian(oracle, abi.F_PLANE, prob.globals, None, params)
        for bi, (Ja, Jnum) in enumerate(zip(J, Jn)):
            ts = Jnum.shape[1]
            assert np.abs(Ja[:, :ts] - Jnum).max() < 1e-3 * max(1.0, np.abs(Jnum).max()), (f, bi)


def test_null_jacobian_blocks_are_skipped(oracle):
    prob, st = window_factor_cases(1)
    params = [block_ptr(st, b) for b in [0, 1, 22, 32, 30]]
    r1, J1 = oracle.factor_evaluate(abi.F_PROJ_2F1C, prob.globals, prob.vis_obs[0], params)
    r2, J2 = oracle.factor_evaluate(abi.F_PROJ_2F1C, prob.globals, prob.vis_obs[0], params, null_jac=(2, 4))
    assert np.array_equal(r1, r2) and J2[2] is None and J2[4] is None
    assert np.array_equal(J1[0], J2[0]) and np.array_equal(J1[3], J2[3])


def test_huber_matches_in_tree_corrector(oracle):
    """Huber(1): s<=1 -> (s,1,0); s>1 -> (2 sqrt(s)-1, 1/sqrt(s), -1/(2 s^1.5)); rho'' < 0 in the outlier branch so
    only the sqrt(rho') scaling is ever taken (marginalization_factor.cpp:46-57)."""
    import ctypes as C
    rho = (C.c_double * 3)()
    oracle.lib().vo_huber(C.c_double(1.0), C.c_double(0.25), rho)
    assert list(rho) == [0.25, 1.0, 0.0]
    oracle.lib().vo_huber(C.c_double(1.0), C.c_double(4.0), rho)
    assert abs(rho[0] - 3.0) < 1e-15 and abs(rho[1] - 0.5) < 1e-15 and abs(rho[2] + 0.0625) < 1e-15


def test_preintegration_matches_generator(oracle):
    """The generator's own numpy pre-integration and the oracle's restatement of integration_base.h /
    wheel_integration_base.h agree (two independent transcriptions)."""
    cfg = synth.make_config(4)
    seq = synth.Sequence(cfg, 0, 11)
    dt, acc, gyr = seq.imu[0]
    ba, bg = np.array([0.01, -0.02, 0.005]), np.array([0.001, 0.002, -0.001])
    noise = (cfg.acc_n, cfg.gyr_n, cfg.acc_w, cfg.gyr_w)
    a = synth.imu_preintegrate(dt, acc, gyr, ba, bg, noise)
    b = oracle.imu_preintegrate(dt, acc, gyr, ba, bg, noise)
    assert np.allclose(a, b, rtol=1e-9, atol=1e-13)
    dt, vel, gyr = seq.wheel[0]
    a = synth.wheel_preintegrate(dt, vel, gyr, (1.01, 0.99, 1.02), 0.001, (cfg.vel_n_wheel, cfg.gyr_n_wheel))
    b = oracle.wheel_preintegrate(dt, vel, gyr, (1.01, 0.99, 1.02), 0.001, (cfg.vel_n_wheel, cfg.gyr_n_wheel))
    assert np.allclose(a, b, rtol=1e-9, atol=1e-13)

"""a-3 (the arithmetic inside ceres::Solve, call site estimator.cpp:1643-1658) from a second, independent angle: the whole Ceres 1.14
TrustRegionMinimizer + DoglegStrategy(TRADITIONAL_DOGLEG) recurrence restated here with DENSE numpy algebra -- no Schur complement, no block-sparse
storage, no code shared with oracle/vo_solver.c beyond the linearisation (normal equations / cost / manifold Plus, which are pinned to the
reference's compiled factor code elsewhere) -- and compared with the oracle's solver TRACE row by row: cost, radius, mu, step norm, Cauchy alpha,
dogleg step norm, model decrease and the valid / successful flags of every iteration, on C1..C4 and C6, on a rejected-step start and on a long run.
Written from the published algorithm (trust_region_minimizer.cc, dogleg_strategy.cc), not from vo_solver.c."""
import numpy as np
import pytest

from viwb import abi, synth
from test_oracle_solver import active_tangent_mask

VIS_TO_T = np.r_[np.arange(66), 165 + np.arange(12), [190]]


def dense_system(oracle, prob, x, act=None):
    """(H, g, cost) over [active fixed tangent columns | landmarks] from the oracle's normal equations at x."""
    H, g, lm, c = oracle.normal_equations(prob, x)
    T, N = abi.TANGENT_FIXED, prob.num_landmarks
    if act is None:
        # Program::RemoveFixedBlocks also drops BLOCKS no residual touches: all their rows / columns are exactly zero (decided once, at the start)
        cols = np.nonzero(active_tangent_mask(prob)[:T])[0]
        keep = np.zeros(T, bool)
        for b in range(abi.NUM_FIXED_BLOCKS):
            o, n_ = abi.block_toffset(b), abi.block_tsize(b)
            blk = [t for t in cols if o <= t < o + n_]
            if blk and (np.any(H[blk][:, blk] != 0.0) or np.any(g[blk] != 0.0)):
                keep[blk] = True
        act = np.nonzero(keep)[0]
    na = len(act)
    n = na + N
    Hf, gf = np.zeros((n, n)), np.zeros(n)
    Hf[:na, :na] = H[np.ix_(act, act)]
    gf[:na] = g[act]
    pos = {t: i for i, t in enumerate(act)}
    cols = np.array([pos.get(t, -1) for t in VIS_TO_T])
    ok = cols >= 0
    for k in range(N):
        Hf[na + k, na + k] = lm[k, 0]
        gf[na + k] = lm[k, 1]
        Hf[cols[ok], na + k] = lm[k, 2:81][ok]
        Hf[na + k, cols[ok]] = lm[k, 2:81][ok]
    return Hf, gf, c, act


def ambient_active(prob):
    m = np.zeros(abi.STATE_FIXED + prob.num_landmarks, bool)
    for b in range(abi.NUM_FIXED_BLOCKS):
        if (prob.block_flags[b] & abi.BLOCK_PRESENT) and not (prob.block_flags[b] & abi.BLOCK_CONSTANT):
            m[abi.block_offset(b): abi.block_offset(b) + abi.block_size(b)] = True
    m[abi.STATE_FIXED:] = True
    return m


def dense_dogleg_minimize(oracle, prob, x0, opt):
    """Ceres 1.14 trust-region loop, dense.  Returns the trace rows (same keys as the oracle's)."""
    min_mu, max_mu, mu_inc = 1e-8, 1.0, 10.0
    T = abi.TANGENT_FIXED
    x = np.array(x0, float)
    H, g, cost, act = dense_system(oracle, prob, x)
    amb = ambient_active(prob)
    # blocks without residuals are not part of the reduced program: leave them out of x_norm like Ceres does
    full_T = np.zeros(T + prob.num_landmarks, bool); full_T[act] = True; full_T[T:] = True
    for b in range(abi.NUM_FIXED_BLOCKS):
        if not full_T[abi.block_toffset(b)]:
            amb[abi.block_offset(b): abi.block_offset(b) + abi.block_size(b)] = False
    scale = 1.0 / (1.0 + np.sqrt(np.diag(H))) if opt.jacobi_scaling else np.ones(len(g))      # fixed at the first linearisation
    radius, mu, reuse = opt.initial_trust_region_radius, min_mu, False
    x_norm = np.linalg.norm(x[amb])
    trace = [dict(iteration=0, cost=cost, radius=radius, mu=mu)]
    it, invalid = 0, 0
    sgrad = gn = D = None
    alpha = 0.0

    def to_full(delta_active):
        d = np.zeros(T + prob.num_landmarks)
        d[act] = delta_active[:len(act)]
        d[T:] = delta_active[len(act):]
        return d
    gmax = np.abs(x - oracle.state_plus(prob, x, to_full(-g)))[amb].max()
    if gmax <= opt.gradient_tolerance:
        return trace
    while True:
        it += 1
        Hs, gs = H * scale[:, None] * scale[None, :], g * scale
        if not reuse:
            D = np.sqrt(np.clip(np.diag(Hs), opt.min_lm_diagonal, opt.max_lm_diagonal))
            sgrad = gs / D
            v = sgrad / D
            alpha = (sgrad @ sgrad) / (v @ Hs @ v)
            ok = False
            while mu < max_mu:
                try:
                    L = np.linalg.cholesky(Hs + mu * np.diag(D * D))
                    y = np.linalg.solve(L.T, np.linalg.solve(L, gs))
                    ok = bool(np.all(np.isfinite(y)))
                except np.linalg.LinAlgError:
                    ok = False
                if ok:
                    break
                mu *= mu_inc
            assert ok, "linear solver failure: not exercised by these windows"
            gn = -D * y
            reuse = True
        gnorm, gnn = np.linalg.norm(sgrad), np.linalg.norm(gn)
        if gnn <= radius:
            step, dsn = gn.copy(), gnn
        elif gnorm * alpha >= radius:
            step, dsn = -(radius / gnorm) * sgrad, radius
        else:
            a_vec = -alpha * sgrad
            b_vec = gn
            bma = b_vec - a_vec
            a2, c = a_vec @ a_vec, a_vec @ bma
            d = np.sqrt(c * c + (bma @ bma) * (radius * radius - a2))
            beta = (d - c) / (bma @ bma) if c <= 0 else (radius * radius - a2) / (d + c)
            step = a_vec + beta * bma
            dsn = np.linalg.norm(step)
        delta_s = step / D                                            # Jacobi-scaled space
        mcc = -(delta_s @ gs) - 0.5 * delta_s @ Hs @ delta_s
        row = dict(iteration=it, alpha=alpha, dogleg_step_norm=dsn, model_cost_change=mcc, mu=mu)
        if not mcc > 0.0:
            invalid += 1
            row.update(step_is_valid=0, step_is_successful=0, cost=cost, radius=radius)
            trace.append(row)
            if invalid >= opt.max_num_consecutive_invalid_steps:
                return trace
            mu *= mu_inc; reuse = False
            if it >= opt.max_num_iterations or radius <= opt.min_trust_region_radius:
                return trace
            continue
        invalid = 0
        cand = oracle.state_plus(prob, x, to_full(delta_s * scale))
        cand_cost = oracle.cost(prob, cand)
        step_norm = np.linalg.norm((x - cand)[amb])
        cost_change = cost - cand_cost
        row.update(step_is_valid=1, step_norm=step_norm, cost_change=cost_change)
        if step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance) or abs(cost_change) <= opt.function_tolerance * cost:
            row.update(step_is_successful=0, cost=cost, radius=radius, converged=1)
            trace.append(row)
            return trace
        rd = cost_change / mcc
        success = rd > opt.min_relative_decrease
        if success:
            x, cost = cand, cand_cost
            H, g, _, _ = dense_system(oracle, prob, x, act)
            x_norm = np.linalg.norm(x[amb])
            if rd < 0.25:
                radius *= 0.5
            if rd > 0.75:
                radius = max(radius, 3.0 * dsn)
            mu = max(min_mu, 2.0 * mu / mu_inc)
            reuse = False
        else:
            radius *= 0.5
            reuse = True
        row.update(step_is_successful=int(success), cost=cost, radius=radius, relative_decrease=rd, mu_after=mu)
        trace.append(row)
        if success:
            gmax = np.abs(x - oracle.state_plus(prob, x, to_full(-g)))[amb].max()
            if gmax <= opt.gradient_tolerance:
                return trace
        if it >= opt.max_num_iterations or radius <= opt.min_trust_region_radius:
            return trace


def compare(tr_ref, tr_np, rel=2e-6):
    assert len(tr_ref) == len(tr_np), (len(tr_ref), len(tr_np))
    assert abs(tr_ref[0]["cost"] - tr_np[0]["cost"]) <= 1e-12 * tr_np[0]["cost"]
    for a, b in zip(tr_ref[1:], tr_np[1:]):
        assert a["iteration"] == b["iteration"] and a["step_is_valid"] == b["step_is_valid"] and a["step_is_successful"] == b["step_is_successful"], (a, b)
        for key in ("alpha", "dogleg_step_norm", "model_cost_change"):
            assert abs(a[key] - b[key]) <= rel * abs(b[key]), (key, a[key], b[key], a["iteration"])
        if b.get("converged"):
            continue
        assert abs(a["radius"] - b["radius"]) <= rel * b["radius"], ("radius", a, b)      # 3 x the dogleg step norm after a very good step: that norm's tolerance
        assert abs(a["cost"] - b["cost"]) <= 1e-7 * b["cost"], ("cost", a["cost"], b["cost"])
        if b["step_is_valid"]:
            assert abs(a["step_norm"] - b["step_norm"]) <= 1e-5 * max(b["step_norm"], 1e-9), ("step_norm", a["step_norm"], b["step_norm"])
            assert abs(a["cost_change"] - b["cost_change"]) <= 1e-6 * abs(b["cost_change"]) + 1e-9 * b["cost"], ("cost_change", a, b)


@pytest.mark.parametrize("cid", [1, 2, 3, 4, 6])
def test_eight_iterations_trace_by_trace(oracle, cid):
    prob, st, _ = synth.make_window(cid)
    opt = abi.default_options()
    sol, summ, tr = oracle.window_solve(prob, st, opt, want_trace=True)
    compare(tr, dense_dogleg_minimize(oracle, prob, st, opt))


def test_window_with_prior_trace_by_trace(oracle):
    seq = synth.Sequence(synth.make_config(2), 3, 12)
    prob, st, _ = seq.window(0)
    a, sm, q = oracle.optimization(prob, st, abi.MARGIN_OLD)
    prob, st, _ = seq.window(1, prior=q, prev_state=a)
    opt = abi.default_options()
    sol, summ, tr = oracle.window_solve(prob, st, opt, want_trace=True)
    compare(tr, dense_dogleg_minimize(oracle, prob, st, opt))


@pytest.mark.parametrize("mult,radius,iters,want_reject", [(6.0, 1e4, 20, True), (3.0, 1.0, 12, False), (3.0, 30.0, 12, False)])
def test_rejected_steps_small_radius_and_long_runs_trace_by_trace(oracle, mult, radius, iters, want_reject):
    """depths far off + a long run: a rejected step (radius halves, the Gauss-Newton step is reused); a small initial radius: the Cauchy and the
    dogleg-interpolation branches with their radius growth"""
    prob, st, _ = synth.make_window(1)
    bad = st.copy()
    bad[abi.STATE_FIXED:] *= mult
    opt = abi.default_options()
    opt.max_num_iterations = iters
    opt.initial_trust_region_radius = radius
    sol, summ, tr = oracle.window_solve(prob, bad, opt, want_trace=True)
    tn = dense_dogleg_minimize(oracle, prob, bad, opt)
    if want_reject:
        assert any(e["step_is_valid"] and not e["step_is_successful"] for e in tr[1:])
    else:
        assert any(abs(e["dogleg_step_norm"] - e["radius"]) < 1e-9 * e["radius"] or e["dogleg_step_norm"] < 0.999 * tr[k]["radius"] for k, e in enumerate(tr[1:]))
    compare(tr, tn, rel=2e-5)

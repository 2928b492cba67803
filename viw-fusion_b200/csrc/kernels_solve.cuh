// kernels_solve.cuh -- one block per window: the Ceres trust-region recurrence (DENSE_SCHUR + DOGLEG) on the
// assembled normal equations.  Restates ceres-solver 1.14 semantics (call site
// estimator.cpp:1643-1658): Jacobi scaling fixed at the first linearisation, D = sqrt(clamp(diag)), Cauchy
// length, mu-regularised Gauss-Newton step through the Schur complement on the inverse depths + dense Cholesky
// (mu x10 on failure), traditional dogleg interpolation, model decrease, candidate = x (+) step; and, one call
// later, the accept/reject decision with the radius / mu updates and all termination tests.
//
// The reduced system (<= 192 active tangent columns) lives in shared memory as a SKYLINE: row i of the lower triangle is stored
// from its first structurally non-zero column efirst[i] (host symbolic analysis, csrc/viwb.cu lower_count).  With the speed-bias
// blocks eliminated first, newest frame first, the stereo+IMU window needs 7.2 K instead of 13.7 K doubles, which lets two
// windows share an SM.  All quadratic forms are evaluated on the *unscaled* matrices with unscaled directions
// u = c o (v / D), which equals Ceres' J_scaled products exactly.
#pragma once
#include "layout.cuh"
#include "factors.cuh"
#include "kernels_asm.cuh"

namespace viwb {

VIWB_HD int pidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// block-wide sum / max; red has >= 32 doubles.  Every thread gets the result (fixed reduction order: deterministic).
#ifdef VIWB_HOST_EMU
VIWB_D double block_sum(double v, int, int, double *) { return v; }
VIWB_D double block_max(double v, int, int, double *) { return v; }
#else
VIWB_D double block_sum(double v, int tid, int nt, double *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    const int nw = nt >> 5;
    double r = 0.0;
    for (int k = 0; k < nw; k++) r += red[k];
    __syncthreads();
    return r;
}
VIWB_D double block_max(double v, int tid, int nt, double *red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    const int nw = nt >> 5;
    double r = red[0];
    for (int k = 1; k < nw; k++) r = fmax(r, red[k]);
    __syncthreads();
    return r;
}
#endif

// several sums in one pass: one barrier pair for all of them (red needs N * 32 doubles)
template <int N>
VIWB_D void block_sum_n(double (&v)[N], int tid, int nt, double *red) {
#ifdef VIWB_HOST_EMU
    (void)v; (void)tid; (void)nt; (void)red;
#else
    for (int q = 0; q < N; q++) for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
    const int nw = nt >> 5;
    if ((tid & 31) == 0) for (int q = 0; q < N; q++) red[q * 32 + (tid >> 5)] = v[q];
    __syncthreads();
    for (int q = 0; q < N; q++) { double r = 0.0; for (int k = 0; k < nw; k++) r += red[q * 32 + k]; v[q] = r; }
    __syncthreads();
#endif
}

struct SolveSmem {
    double *L;       // skyline lower triangle: row i holds columns fst[i]..i at L[rp[i] .. rp[i+1])
    int *rp, *fst;   // row pointers (nf + 1) and first columns (nf)
    double *g, *sc, *D, *sg, *y, *u, *Hu, *ug, *uvis, *red, *bc, *chol, *dinv, *xo, *Pt;
    int *amap, *vmap;
};
VIWB_HD size_t solve_smem_doubles(int nt, int esize) { return (size_t)((esize + 1) & ~1) + 9 * TFIX + 2 * VSUB + 5 * 32 + 16 + 2 * TFIX + 2 + (size_t)((nt + 31) / 32) * 48 + 2 * TFIX + 8 * (TFIX + 4) + 2; }   // ints: amap, vmap, rp, fst
VIWB_D void carve(SolveSmem &s, double *smem, int nt, int esize) {
    double *p = smem;
    s.L = p; p += (size_t)((esize + 1) & ~1);
    s.g = p; p += TFIX; s.sc = p; p += TFIX; s.D = p; p += TFIX; s.sg = p; p += TFIX; s.y = p; p += TFIX;
    s.u = p; p += TFIX; s.Hu = p; p += TFIX; s.ug = p; p += TFIX;
    s.uvis = p; p += 2 * VSUB;
    s.red = p; p += 5 * 32; s.bc = p; p += 16;
    s.chol = p; p += (size_t)((nt + 31) / 32) * 48;
    s.dinv = p; p += TFIX; s.xo = p; p += TFIX;
    if ((p - smem) & 1) p++;                    // Pt is read with 16-byte vector loads
    s.Pt = p; p += 8 * (TFIX + 4);
    s.amap = (int *)p; s.vmap = s.amap + TFIX; s.rp = s.vmap + TFIX; s.fst = s.rp + TFIX + 2;
}
// entry (i, j), j <= i, of the skyline; must lie inside the envelope
#define SKY(s, i, j) ((s).L[(s).rp[i] - (s).fst[i] + (j)])

// assemble the active part of H (unscaled) and g into the skyline / s.g, and keep a copy in HBM
VIWB_D void assemble_H(const BatchDev &bd, int w, const SolveSmem &s, int nf, double *Hpk, double *gpk, int tid, int nt) {
    const int ne = s.rp[nf];
    for (int e = tid; e < ne; e += nt) s.L[e] = 0.0;
    for (int i = tid; i < nf; i += nt) s.g[i] = 0.0;
    VIWB_SYNC();
    PackedTarget t; t.L = s.L; t.g = s.g; t.tcol = bd.meta[w].tcol; t.rp = s.rp; t.fst = s.fst;
    assemble_into(t, bd, w, MODE_SOLVE, tid, nt, (int *)s.Hu);      // Hu is free until the step computation
    for (int e = tid; e < ne; e += nt) Hpk[e] = s.L[e];
    for (int i = tid; i < nf; i += nt) gpk[i] = s.g[i];
    VIWB_SYNC();
}
// reload the copy (after the Cholesky factor overwrote it)
VIWB_D void load_H(const double *Hpk, const SolveSmem &s, int nf, int tid, int nt) {
    const int ne = s.rp[nf];
    for (int e = tid; e < ne; e += nt) s.L[e] = Hpk[e];
    VIWB_SYNC();
}
// Hu = H u over the skyline (u, Hu of length nf): the stored part of row i, then column i of the rows below that reach it
VIWB_D void symv(const SolveSmem &s, int nf, const double *u, double *Hu, int tid, int nt) {
    for (int i = tid; i < nf; i += nt) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;            // four partial sums (a dependent FMA chain of up to nf otherwise)
        const double *ri = s.L + s.rp[i] - s.fst[i];
        int j = s.fst[i];
        for (; j + 3 <= i; j += 4) { a0 += ri[j] * u[j]; a1 += ri[j + 1] * u[j + 1]; a2 += ri[j + 2] * u[j + 2]; a3 += ri[j + 3] * u[j + 3]; }
        for (; j <= i; j++) a0 += ri[j] * u[j];
        int k = i + 1;
        for (; k + 1 < nf; k += 2) {
            if (s.fst[k] <= i) a1 += s.L[s.rp[k] - s.fst[k] + i] * u[k];
            if (s.fst[k + 1] <= i) a2 += s.L[s.rp[k + 1] - s.fst[k + 1] + i] * u[k + 1];
        }
        for (; k < nf; k++) if (s.fst[k] <= i) a3 += s.L[s.rp[k] - s.fst[k] + i] * u[k];
        Hu[i] = (a0 + a1) + (a2 + a3);
    }
    VIWB_SYNC();
}
// uvis[p] = u mapped to the visual subspace (0 where the column is inactive)
VIWB_D void to_vis(const SolveSmem &s, int nf, const double *u, double *uvis, int tid, int nt) {
    for (int p = tid; p < VSUB; p += nt) uvis[p] = 0.0;
    VIWB_SYNC();
    for (int i = tid; i < nf; i += nt) if (s.vmap[i] >= 0) uvis[s.vmap[i]] = u[i];
    VIWB_SYNC();
}
VIWB_D double dot80(const double *W, const double *v) { double a = 0.0; for (int p = 0; p < 79; p++) a += W[p] * v[p]; return a; }
// the same product by one warp: coalesced reads of the landmark's W row, every lane gets the sum (fixed order: deterministic)
VIWB_D double warp_dot80(const double *W, const double *v, int lane) {
#ifdef VIWB_HOST_EMU
    (void)lane; return dot80(W, v);
#else
    double a = W[lane] * v[lane] + W[lane + 32] * v[lane + 32];
    if (lane < 15) a += W[lane + 64] * v[lane + 64];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
#endif
}
// landmark loops run one warp per landmark: lane 0 carries the scalar work and the partial sums
#ifdef VIWB_HOST_EMU
#define VIWB_LM_LOOP(k, N) for (int k = 0, lane = 0; k < (N); k++)
#else
#define VIWB_LM_LOOP(k, N) for (int k = tid >> 5, lane = tid & 31; k < (N); k += nt >> 5)
#endif

// Blocked (panel width 8) in-place Cholesky of the packed lower triangle with the right-hand side carried as an extra
// row n, so that on exit y = L^-1 b (forward substitution for free).  Returns false on a non-positive pivot
// (Eigen LLT semantics: schur_complement_solver.cc -> LINEAR_SOLVER_FAILURE).  2 barriers per panel:
//   every warp factors the 8x8 diagonal block redundantly (8 lanes, one row each, in registers; pivots through rsqrt so
//   that the panel rows multiply instead of divide) into its own scratch, so no barrier separates it from the panel solve;
//   one thread per panel row, which also drops its 8 results into a column-major panel buffer Pt[8][rows]; the trailing
//   update then runs 4x4 register tiles whose operands are 32-byte vector loads from Pt (conflict-free), not strided
//   reads of the packed triangle.
// scratch: nwarps * 48 doubles (36 packed block entries + 8 inverse pivots + flag); dinv[n] receives 1 / L_kk;
// Pt: 8 * CHOL_LDP doubles, 16-byte aligned.
enum { CHOL_NB = 8, CHOL_SCR = 48, CHOL_LDP = TFIX + 4 };
VIWB_D void load4(const double *p, double *o) {
#ifdef VIWB_HOST_EMU
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3];
#else
    const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
#endif
}
VIWB_D bool cholesky_packed_rhs(double *L, const int *rp, const int *fst, double *y, int n, int tid, int nt, double *scratch, double *dinv, double *Pt) {
    // skyline rows: ROW(i)[j] is entry (i, j) for fst[i] <= j <= i; row n is the right-hand side (dense).  FST(i) = first stored column.
#define ROW(i) ((i) < n ? L + rp[i] - fst[i] : y)
#define FST(i) ((i) < n ? fst[i] : 0)
    const int lane = tid & 31;
    double *my = scratch + (size_t)(tid >> 5) * CHOL_SCR;      // this warp's copy of the factored diagonal block
    for (int c0 = 0; c0 < n; c0 += CHOL_NB) {
        const int nb = (n - c0) < CHOL_NB ? (n - c0) : CHOL_NB;
#ifdef VIWB_HOST_EMU
        {   // single-thread statement of the same arithmetic
            (void)lane;
            my[44] = 1.0;
            double a[8][8];
            for (int i = 0; i < nb; i++) for (int j = 0; j <= i; j++) a[i][j] = (c0 + j >= FST(c0 + i)) ? ROW(c0 + i)[c0 + j] : 0.0;
            for (int k = 0; k < nb; k++) {
                const double d = a[k][k];
                if (!(d > 0.0)) { my[44] = 0.0; break; }
                const double inv = 1.0 / sqrt(d);
                a[k][k] = d * inv; my[36 + k] = inv;
                for (int i = k + 1; i < nb; i++) a[i][k] *= inv;
                for (int i = k + 1; i < nb; i++) for (int j = k + 1; j <= i; j++) a[i][j] -= a[i][k] * a[j][k];
            }
            for (int i = 0; i < nb; i++) for (int j = 0; j <= i; j++) my[i * (i + 1) / 2 + j] = a[i][j];
        }
#else
        {
            double a[8];
            const bool act = lane < nb;
            const int rr = c0 + (act ? lane : 0), fr = fst[rr];
            const double *src = ROW(rr) + c0;
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = (act && j <= lane && c0 + j >= fr) ? src[j] : 0.0;
            bool okp = true;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double d = __shfl_sync(0xffffffffu, a[k], k);
                if (k < nb && !(d > 0.0)) okp = false;
                const double inv = (k < nb && d > 0.0) ? rsqrt(d) : 0.0;
                if (lane == k) { a[k] = d * inv; if (k < nb) my[36 + k] = inv; }
                else if (lane > k) a[k] *= inv;
#pragma unroll
                for (int j = k + 1; j < 8; j++) {
                    const double ljk = __shfl_sync(0xffffffffu, a[k], j);      // L[j][k]
                    if (lane >= j) a[j] -= a[k] * ljk;
                }
            }
            if (act) {
#pragma unroll
                for (int j = 0; j < 8; j++) if (j <= lane) my[lane * (lane + 1) / 2 + j] = a[j];
            }
            if (lane == 0) my[44] = okp ? 1.0 : 0.0;
            __syncwarp();
        }
#endif
        if (my[44] == 0.0) return false;                    // every warp reaches the same verdict
        const int r0 = c0 + nb;
        const int nr = n + 1 - r0, mt = (nr + 3) / 4;       // panel rows r0..n (row n = rhs), padded to whole 4-row tiles
        // panel rows: x_k = (a_ik - sum_m x_m L_km) / L_kk; columns left of a row's envelope are (and stay) structural zeros
        for (int i = r0 + tid; i < r0 + 4 * mt; i += nt) {
            double x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = 0.0;
            if (i <= n && FST(i) < r0) {
                double *ri = ROW(i) + c0;
                const int fi = FST(i);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (k < nb && c0 + k >= fi) {
                        double v = ri[k];
#pragma unroll
                        for (int m = 0; m < 8; m++) if (m < k) v -= x[m] * my[k * (k + 1) / 2 + m];
                        x[k] = v * my[36 + k];
                    }
                }
                // (stored after the chain: a store to the row in between would pin every later load of the diagonal block behind it)
#pragma unroll
                for (int k = 0; k < 8; k++) if (k < nb && c0 + k >= fi) ri[k] = x[k];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) Pt[k * CHOL_LDP + (i - r0)] = x[k];
        }
        for (int r = tid; r < nb; r += nt) dinv[c0 + r] = my[36 + r];
        VIWB_SYNC();
        // factored diagonal block back into the matrix (only now: the other warps have finished reading the unfactored one)
        for (int r = tid; r < nb; r += nt) { double *rk = ROW(c0 + r) + c0; const int fr = FST(c0 + r); for (int j = 0; j <= r; j++) if (c0 + j >= fr) rk[j] = my[r * (r + 1) / 2 + j]; }
        // trailing update with 4x4 register tiles over rows i in [r0, n], columns k in [r0, min(i, n-1)]; a tile whose row group or
        // column group lies entirely right of the panel (first stored column >= r0) has nothing to receive
        const int ntile = mt * (mt + 1) / 2;
        for (int t = tid; t < ntile; t += nt) {
            int ti, tk; sym_unrank(t, ti, tk);
            const int ib = r0 + 4 * ti, kb = r0 + 4 * tk;
            bool hit_i = false, hit_k = false;
#pragma unroll
            for (int a = 0; a < 4; a++) { if (ib + a <= n && FST(ib + a) < r0) hit_i = true; if (kb + a < n && FST(kb + a) < r0) hit_k = true; }
            if (!hit_i || !hit_k) continue;
            double acc[16];
#pragma unroll
            for (int q = 0; q < 16; q++) acc[q] = 0.0;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                double li[4], lk[4];
                load4(Pt + m * CHOL_LDP + 4 * ti, li); load4(Pt + m * CHOL_LDP + 4 * tk, lk);
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b2 = 0; b2 < 4; b2++) acc[a * 4 + b2] += li[a] * lk[b2];
            }
#ifndef SOLVE_NO_BATCH_RMW
            // all sixteen current values are fetched before the first store (four different row pointers: the compiler has to assume they alias)
            {
                double *rp4[4]; bool rowok[4], colok[4];
#pragma unroll
                for (int a = 0; a < 4; a++) { const int i = ib + a; rowok[a] = i <= n && FST(i) < r0; rp4[a] = rowok[a] ? ROW(i) : L; }
#pragma unroll
                for (int b2 = 0; b2 < 4; b2++) { const int k = kb + b2; colok[b2] = k < n && FST(k) < r0; }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b2 = 0; b2 < 4; b2++) { const bool on = rowok[a] && colok[b2] && kb + b2 <= ib + a; acc[a * 4 + b2] = on ? rp4[a][kb + b2] - acc[a * 4 + b2] : 0.0; }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b2 = 0; b2 < 4; b2++) { if (rowok[a] && colok[b2] && kb + b2 <= ib + a) rp4[a][kb + b2] = acc[a * 4 + b2]; }
            }
#else
#pragma unroll
            for (int a = 0; a < 4; a++) { const int i = ib + a; if (i <= n && FST(i) < r0) { double *ri = ROW(i);
#pragma unroll
                for (int b2 = 0; b2 < 4; b2++) { const int k = kb + b2; if (k < n && k <= i && FST(k) < r0) ri[k] -= acc[a * 4 + b2]; } } }
#endif
        }
        VIWB_SYNC();
    }
#undef ROW
#undef FST
    return true;
}
// back substitution L^T x = y (blocks of 8 from the bottom): every thread solves the 8x8 triangular block redundantly from
// broadcast reads (no barrier between the block solve and the update), then all threads update the rows above.  One barrier
// per block.  Result in y; xo is scratch of n doubles.
VIWB_D void chol_backsolve_blocked(const double *L, const int *rp, const int *fst, double *y, const double *dinv, double *xo, int n, int tid, int nt) {
    const int nblk = (n + CHOL_NB - 1) / CHOL_NB;
    for (int bi = nblk - 1; bi >= 0; bi--) {
        const int c0 = bi * CHOL_NB, nb = (n - c0) < CHOL_NB ? (n - c0) : CHOL_NB;
        double x[8];
        const double *rw[8]; int fr[8];              // the block's skyline rows and their first columns
#pragma unroll
        for (int m = 0; m < 8; m++) { const int r = (m < nb) ? c0 + m : c0; rw[m] = L + rp[r] - fst[r]; fr[m] = fst[r]; }
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            x[k] = 0.0;
            if (k < nb) {
                double v = y[c0 + k];
#pragma unroll
                for (int m = 7; m > k; m--) if (m < nb && c0 + k >= fr[m]) v -= rw[m][c0 + k] * x[m];
                x[k] = v * dinv[c0 + k];
            }
        }
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < nb) xo[c0 + k] = x[k];
        }
        for (int i = tid; i < c0; i += nt) {
            double v = y[i];
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < nb && i >= fr[k]) v -= rw[k][i] * x[k];
            y[i] = v;
        }
        VIWB_SYNC();
    }
    for (int i = tid; i < n; i += nt) y[i] = xo[i];
    VIWB_SYNC();
}

VIWB_D void terminate(WinWork &ww, int term) { ww.status = ST_DONE; ww.term = term; }

// max |x - Plus(x, -g)| over one fixed block (ambient infinity norm of the projected gradient step)
VIWB_D double block_grad_inf(int b, unsigned mask, const double *x, const double *gneg) {
    const int gs = blk_size(b);
    double out[9], m = 0.0;          // the largest fixed block is a speed-bias (9)
    if (gs == 7) pose_plus(x, gneg, mask, out);
    else if (gs == 4) quat_plus(x, gneg, mask, out);
    else for (int i = 0; i < gs; i++) out[i] = x[i] + gneg[i];
    for (int i = 0; i < gs; i++) m = fmax(m, fabs(x[i] - out[i]));
    return m;
}

VIWB_D void solve_block(const BatchDev &bd, int bx, int by, int tid, int nt, double *smem, int mode) {
    (void)by;
    const int w = bx;
    const WinMeta &m = bd.meta[w];
    WinWork &ww = bd.work[w];
    if (ww.status != ST_RUNNING) return;
    const Opts &op = bd.opt;
    SolveSmem s; carve(s, smem, nt, bd.env_max);
    const int nf = m.nf, N = m.nlm, vo = vec_off(bd, w);
    for (int i = tid; i < nf; i += nt) s.fst[i] = m.efirst[i];
    if (tid == 0) { int acc = 0; for (int i = 0; i < nf; i++) { s.rp[i] = acc; acc += i - m.efirst[i] + 1; } s.rp[nf] = acc; }
    double *x = bd.x_cur + m.state_off, *xc = bd.x_cand + m.state_off;
    double *g_scale = bd.v_scale + vo, *g_D = bd.v_D + vo, *g_sg = bd.v_sgrad + vo, *g_gn = bd.v_gn + vo;
    double *wug = bd.v_wug + m.lm_off, *wun = bd.v_wun + m.lm_off;      // per landmark: w_k . u_g (Cauchy direction) and w_k . (C y) (Gauss-Newton step), reused by the quadratic forms
    const double *lm_a = bd.lm_a + m.lm_off, *lm_g = bd.lm_g + m.lm_off, *lm_gamma = bd.lm_gamma + m.lm_off, *lm_sc = bd.lm_scale + m.lm_off;
    const double *W = bd.lm_W + (size_t)m.lm_off * VSUB;
    double *Hpk = bd.Hpk + (size_t)w * (TFIX * (TFIX + 1) / 2), *gpk = bd.gpk + (size_t)w * TFIX;
    const double *Tvis = bd.Tvis + (size_t)w * VSUB * VSUB, *tvec = bd.tvec + (size_t)w * VSUB;
    const double min_mu = 1e-8, max_mu = 1.0, mu_inc = 10.0;

    // compact <-> fixed / visual index maps
    for (int b = tid; b < NB; b += nt) if (m.tcol[b] >= 0) for (int k = 0; k < blk_tsize(b); k++) {
        s.amap[m.tcol[b] + k] = blk_toff(b) + k;
        s.vmap[m.tcol[b] + k] = blk_voff(b) >= 0 ? blk_voff(b) + k : -1;
    }
    VIWB_SYNC();

    if (mode == 2) { assemble_H(bd, w, s, nf, Hpk, gpk, tid, nt); return; }     // debug hook: normal equations only
    bool assembled = false;

    // ================= phase A: decision on the pending candidate ==================================
    double c_part = 0.0;
    for (int k = tid; k < N; k += nt) c_part += bd.lm_cost[m.lm_off + k];
    const double cand_cost = block_sum(c_part, tid, nt, s.red) + ww.small_cost;
    bool new_lin = false;
    if (ww.phase == PH_INIT) {
        if (tid == 0) { ww.x_cost = cand_cost; ww.initial_cost = cand_cost; ww.num_iterations = 1; }
        new_lin = true;
        for (int i = tid; i < SFIX + N; i += nt) x[i] = xc[i];
    } else {
        // ParameterToleranceReached / FunctionToleranceReached / IsStepSuccessful (trust_region_minimizer.cc)
        int verdict = 0;   // 0 rejected, 1 accepted, 2 converged
        const double cost_change = ww.x_cost - cand_cost;
        if (ww.step_norm <= op.parameter_tolerance * (ww.x_norm + op.parameter_tolerance)) verdict = 2;
        else if (fabs(cost_change) <= op.function_tolerance * ww.x_cost) verdict = 2;
        else if (cost_change / ww.model_cost_change > op.min_relative_decrease) verdict = 1;
        if (verdict == 2) { VIWB_SYNC(); if (tid == 0) terminate(ww, 0); return; }
        if (verdict == 1) {
            for (int i = tid; i < SFIX + N; i += nt) x[i] = xc[i];
            new_lin = true;
        }
        VIWB_SYNC();
        if (tid == 0) {
            const double rd = cost_change / ww.model_cost_change;
            if (verdict == 1) {
                ww.x_cost = cand_cost; ww.successful++;
                if (rd < 0.25) ww.radius *= 0.5;                                 // DoglegStrategy::StepAccepted
                if (rd > 0.75) ww.radius = fmax(ww.radius, 3.0 * ww.dogleg_step_norm);
                ww.mu = fmax(min_mu, 2.0 * ww.mu / mu_inc);
                ww.reuse = 0;
            } else { ww.radius *= 0.5; ww.reuse = 1; }                           // StepRejected
            ww.num_iterations++;
        }
        VIWB_SYNC();
        // FinalizeIterationAndCheckIfMinimizerCanContinue (gradient tolerance is tested in phase B)
        if (ww.iteration >= op.max_num_iterations || mode == 3) { VIWB_SYNC(); if (tid == 0) terminate(ww, 1); return; }   // mode 3: max_solver_time reached
        if (ww.radius <= op.min_radius) { VIWB_SYNC(); if (tid == 0) terminate(ww, 0); return; }
    }
    VIWB_SYNC();
    if (new_lin) {   // x_norm over the active blocks
        double a = 0.0;
        for (int b = tid; b < NB; b += nt) if (m.tcol[b] >= 0) for (int k = 0; k < blk_size(b); k++) a += x[blk_off(b) + k] * x[blk_off(b) + k];
        for (int k = tid; k < N; k += nt) a += x[SFIX + k] * x[SFIX + k];
        a = block_sum(a, tid, nt, s.red);
        if (tid == 0) ww.x_norm = sqrt(a);
    }
    if (ww.phase == PH_INIT && op.max_num_iterations <= 0) { VIWB_SYNC(); if (tid == 0) terminate(ww, 1); return; }
    VIWB_SYNC();

    // ================= phase B: compute a trust-region step (loops while the step is invalid) =========
    for (;;) {
        if (tid == 0) ww.iteration++;
        VIWB_SYNC();
        bool ls_failure = false;
        if (!ww.reuse) {
            VIWB_SYNC();
            if (tid == 0) ww.reuse = 1;
            if (!assembled) { assemble_H(bd, w, s, nf, Hpk, gpk, tid, nt); assembled = true; }
            else { load_H(Hpk, s, nf, tid, nt); for (int i = tid; i < nf; i += nt) s.g[i] = gpk[i]; VIWB_SYNC(); }
            if (ww.first) {
                for (int i = tid; i < nf; i += nt) g_scale[i] = op.jacobi_scaling ? 1.0 / (1.0 + sqrt(SKY(s, i, i))) : 1.0;
            }
            VIWB_SYNC();
            for (int i = tid; i < nf; i += nt) s.sc[i] = g_scale[i];
            VIWB_SYNC();
            if (new_lin) {
                // gradient_max_norm = |x - Plus(x, -g)|_inf  (unscaled gradient)
                double gm = 0.0;
                for (int b = tid; b < NB; b += nt) if (m.tcol[b] >= 0) {
                    double gneg[9];
                    for (int k = 0; k < blk_tsize(b); k++) gneg[k] = -s.g[m.tcol[b] + k];
                    gm = fmax(gm, block_grad_inf(b, m.mask[b], x + blk_off(b), gneg));
                }
                for (int k = tid; k < N; k += nt) gm = fmax(gm, fabs(lm_g[k]));
                gm = block_max(gm, tid, nt, s.red);
                if (tid == 0) ww.gradient_max_norm = gm;
                if (gm <= op.gradient_tolerance) { VIWB_SYNC(); if (tid == 0) terminate(ww, 0); return; }
            }
            // D = sqrt(clamp(diag(J_s^T J_s))), scaled gradient, Cauchy direction u_g = c o (sgrad / D)
            double n2 = 0.0;
            for (int i = tid; i < nf; i += nt) {
                double d2 = s.sc[i] * s.sc[i] * SKY(s, i, i);
                d2 = fmin(fmax(d2, op.min_lm_diagonal), op.max_lm_diagonal);
                const double D = sqrt(d2); s.D[i] = D; g_D[i] = D;
                const double sg = s.sc[i] * s.g[i] / D; s.sg[i] = sg; g_sg[i] = sg; n2 += sg * sg;
                s.ug[i] = s.sc[i] * sg / D;
            }
            for (int k = tid; k < N; k += nt) {
                const double c = lm_sc[k];
                double d2 = c * c * lm_a[k]; d2 = fmin(fmax(d2, op.min_lm_diagonal), op.max_lm_diagonal);
                const double D = sqrt(d2); g_D[TFIX + k] = D;
                const double sg = c * lm_g[k] / D; g_sg[TFIX + k] = sg; n2 += sg * sg;
            }
            n2 = block_sum(n2, tid, nt, s.red);
            // q_gg = |J u_g|^2, l_g = g . u_g
            symv(s, nf, s.ug, s.Hu, tid, nt);
            to_vis(s, nf, s.ug, s.uvis, tid, nt);
            double q = 0.0, l = 0.0;
            for (int i = tid; i < nf; i += nt) { q += s.ug[i] * s.Hu[i]; l += s.g[i] * s.ug[i]; }
            VIWB_LM_LOOP(k, N) {
                // the landmark's scalars are fetched (warp-uniform loads) before the dot product so that both latencies overlap
                const double lsc = lm_sc[k], lsg = g_sg[TFIX + k], lD = g_D[TFIX + k], la = lm_a[k], lg = lm_g[k];
                const double dw = warp_dot80(W + (size_t)k * VSUB, s.uvis, lane);
                if (lane == 0) {
                    wug[k] = dw;
                    const double ul = lsc * lsg / lD;
                    q += 2.0 * ul * dw + la * ul * ul;
                    l += lg * ul;
                }
            }
            { double v2[2] = {q, l}; block_sum_n<2>(v2, tid, nt, s.red); q = v2[0]; l = v2[1]; }
            if (tid == 0) { ww.sgrad_norm = sqrt(n2); ww.q_gg = q; ww.l_g = l; ww.alpha = n2 / q; }
            VIWB_SYNC();
            // ---- Gauss-Newton step: (J^T J + mu D^2) y = J^T r through the Schur complement
            ls_failure = true;
            bool h_dirty = false;
            while (ww.mu < max_mu) {
                if (h_dirty) load_H(Hpk, s, nf, tid, nt);
                const double mu = ww.mu;
                const bool t_ok = (mu == ww.mu_lin);          // gamma / T / tvec were built for mu_lin
                // S = C (H - T) C + mu D^2 ; rhs = C (g - tvec)     (one warp per skyline row, lanes over its stored columns)
                {
                    const int Wd = nt < 32 ? nt : 32, nwp = nt / Wd, wid = tid / Wd, ln = tid % Wd;
                    for (int i = wid; i < nf; i += nwp) {
                        const int vi = s.vmap[i];
                        double *ri = s.L + s.rp[i] - s.fst[i];
                        for (int j = s.fst[i] + ln; j <= i; j += Wd) {
                            const int vj = s.vmap[j];
                            double hij = ri[j];
                            if (vi >= 0 && vj >= 0) {
                                if (t_ok) hij -= Tvis[vi * VSUB + vj];
                                else {
                                    double t = 0.0;
                                    for (int k = 0; k < N; k++) {
                                        const double c = lm_sc[k], sk = c * c * lm_a[k], Dk = g_D[TFIX + k];
                                        t += (c * c / (sk + mu * Dk * Dk)) * W[(size_t)k * VSUB + vi] * W[(size_t)k * VSUB + vj];
                                    }
                                    hij -= t;
                                }
                            }
                            hij *= s.sc[i] * s.sc[j];
                            if (i == j) hij += mu * s.D[i] * s.D[i];
                            ri[j] = hij;
                        }
                    }
                    for (int i = tid; i < nf; i += nt) {
                        const int vi = s.vmap[i];
                        double r = s.g[i];
                        if (vi >= 0) {
                            if (t_ok) r -= tvec[vi];
                            else { double t = 0.0; for (int k = 0; k < N; k++) { const double c = lm_sc[k], sk = c * c * lm_a[k], Dk = g_D[TFIX + k]; t += (c * c / (sk + mu * Dk * Dk)) * W[(size_t)k * VSUB + vi] * lm_g[k]; } r -= t; }
                        }
                        s.y[i] = s.sc[i] * r;
                    }
                }
                VIWB_SYNC();
                h_dirty = true;
                if (tid == 0) ww.num_linear++;
                bool ok = cholesky_packed_rhs(s.L, s.rp, s.fst, s.y, nf, tid, nt, s.chol, s.dinv, s.Pt);
                if (ok) {
                    chol_backsolve_blocked(s.L, s.rp, s.fst, s.y, s.dinv, s.xo, nf, tid, nt);
                    // back-substitute the inverse depths (scaled): y_k = (c g_k - c w_k . (C y_f)) / h_k
                    for (int i = tid; i < nf; i += nt) s.u[i] = s.sc[i] * s.y[i];
                    VIWB_SYNC();
                    to_vis(s, nf, s.u, s.uvis, tid, nt);
                    double bad = 0.0;
                    for (int i = tid; i < nf; i += nt) { if (!isfinite(s.y[i])) bad = 1.0; g_gn[i] = -s.D[i] * s.y[i]; }
                    VIWB_LM_LOOP(k, N) {
                        const double c = lm_sc[k], Dk = g_D[TFIX + k], la = lm_a[k], lg = lm_g[k];
                        const double dw = warp_dot80(W + (size_t)k * VSUB, s.uvis, lane);
                        if (lane == 0) {
                            wun[k] = dw;
                            const double hk = c * c * la + mu * Dk * Dk;
                            const double yk = c * (lg - dw) / hk;
                            if (!isfinite(yk)) bad = 1.0;
                            g_gn[TFIX + k] = -Dk * yk;
                        }
                    }
                    bad = block_sum(bad, tid, nt, s.red);
                    ok = (bad == 0.0);
                }
                if (!ok) { VIWB_SYNC(); if (tid == 0) ww.mu *= mu_inc; VIWB_SYNC(); continue; }
                ls_failure = false;
                break;
            }
            VIWB_SYNC();
            if (!ls_failure) {
                // q_gn, q_nn, l_n with u_n = c o (gn / D) = -c o y: s.u still holds c o y from the back-substitution, and the landmark
                // rows' products with it (wun) and with u_g (wug) are on file -- no third pass over W
                load_H(Hpk, s, nf, tid, nt);
                for (int i = tid; i < nf; i += nt) s.u[i] = -s.u[i];
                VIWB_SYNC();
                symv(s, nf, s.u, s.Hu, tid, nt);
                double qnn = 0.0, qgn = 0.0, ln = 0.0, nn = 0.0, gd = 0.0;
                for (int i = tid; i < nf; i += nt) {
                    qnn += s.u[i] * s.Hu[i]; qgn += s.ug[i] * s.Hu[i]; ln += s.g[i] * s.u[i];
                    nn += g_gn[i] * g_gn[i]; gd += s.sg[i] * g_gn[i];
                }
                for (int k = tid; k < N; k += nt) {
                    const double c = lm_sc[k], Dk = g_D[TFIX + k], ggn = g_gn[TFIX + k], gsg = g_sg[TFIX + k], la = lm_a[k], lg = lm_g[k];
                    const double dn = -wun[k], dg = wug[k];
                    const double un = c * ggn / Dk, ug = c * gsg / Dk;
                    qnn += 2.0 * un * dn + la * un * un;
                    qgn += ug * dn + un * dg + la * ug * un;
                    ln += lg * un;
                    nn += ggn * ggn; gd += gsg * ggn;
                }
                { double v5[5] = {qnn, qgn, ln, nn, gd}; block_sum_n<5>(v5, tid, nt, s.red); qnn = v5[0]; qgn = v5[1]; ln = v5[2]; nn = v5[3]; gd = v5[4]; }
                if (tid == 0) { ww.q_nn = qnn; ww.q_gn = qgn; ww.l_n = ln; ww.gn_norm = sqrt(nn); ww.sgrad_dot_gn = gd; }
            }
            VIWB_SYNC();
            if (tid == 0) ww.first = 0;
            VIWB_SYNC();
        }
        // ---- ComputeTraditionalDoglegStep: step = ca * sgrad + cb * gn (scaled space), then / D
        double ca = 0.0, cb = 0.0, mcc = 0.0, dsn = 0.0;
        bool valid = false;
        if (!ls_failure) {
            const double gnorm = ww.sgrad_norm, gnn = ww.gn_norm, radius = ww.radius, alpha = ww.alpha;
            if (gnn <= radius) { ca = 0.0; cb = 1.0; dsn = gnn; }
            else if (gnorm * alpha >= radius) { ca = -(radius / gnorm); cb = 0.0; dsn = radius; }
            else {
                const double b_dot_a = -alpha * ww.sgrad_dot_gn, a_sq = (alpha * gnorm) * (alpha * gnorm);
                const double bma = a_sq - 2.0 * b_dot_a + gnn * gnn, c = b_dot_a - a_sq;
                const double d = sqrt(c * c + bma * (radius * radius - a_sq));
                const double beta = (c <= 0) ? (d - c) / bma : (radius * radius - a_sq) / (d + c);
                ca = -alpha * (1.0 - beta); cb = beta;
                dsn = sqrt(ca * ca * gnorm * gnorm + 2.0 * ca * cb * ww.sgrad_dot_gn + cb * cb * gnn * gnn);
            }
            // model_cost_change = -(J t)^T (r + J t / 2) = -(ca l_g + cb l_n) - (ca^2 q_gg + 2 ca cb q_gn + cb^2 q_nn) / 2
            mcc = -(ca * ww.l_g + cb * ww.l_n) - 0.5 * (ca * ca * ww.q_gg + 2.0 * ca * cb * ww.q_gn + cb * cb * ww.q_nn);
            valid = mcc > 0.0;
        }
        VIWB_SYNC();
        if (!valid) {
            // HandleInvalidStep + DoglegStrategy::StepIsInvalid
            bool stop = false;
            if (tid == 0) {
                ww.num_invalid++;
                if (ww.num_invalid >= op.max_invalid) { terminate(ww, 2); }
                else {
                    ww.mu *= mu_inc; ww.reuse = 0; ww.num_iterations++;
                    if (ww.iteration >= op.max_num_iterations) terminate(ww, 1);
                    else if (ww.radius <= op.min_radius) terminate(ww, 0);
                }
            }
            VIWB_SYNC();
            stop = (ww.status != ST_RUNNING);
            if (stop) return;
            new_lin = false;
            continue;
        }
        // ---- candidate: x (+) (step o scale)
        double sn = 0.0;
        for (int b = tid; b < NB; b += nt) {
            const int o = blk_off(b), gs = blk_size(b);
            if (m.tcol[b] >= 0) {
                double d[9], out[9];
                for (int k = 0; k < blk_tsize(b); k++) { const int i = m.tcol[b] + k; d[k] = (ca * g_sg[i] + cb * g_gn[i]) / g_D[i] * g_scale[i]; }
                if (gs == 7) pose_plus(x + o, d, m.mask[b], out);
                else if (gs == 4) quat_plus(x + o, d, m.mask[b], out);
                else for (int k = 0; k < gs; k++) out[k] = x[o + k] + d[k];
                for (int k = 0; k < gs; k++) { xc[o + k] = out[k]; const double e = x[o + k] - out[k]; sn += e * e; }
            } else for (int k = 0; k < gs; k++) xc[o + k] = x[o + k];
        }
        for (int k = tid; k < N; k += nt) {
            const double d = (ca * g_sg[TFIX + k] + cb * g_gn[TFIX + k]) / g_D[TFIX + k] * lm_sc[k];
            xc[SFIX + k] = x[SFIX + k] + d; sn += d * d;
        }
        sn = block_sum(sn, tid, nt, s.red);
        if (tid == 0) {
            ww.num_invalid = 0; ww.step_norm = sqrt(sn); ww.model_cost_change = mcc; ww.dogleg_step_norm = dsn;
            ww.phase = PH_CAND;
            ww.mu_lin = fmax(min_mu, 2.0 * ww.mu / mu_inc);      // the candidate's linearisation is solved with the post-accept mu
        }
        return;
    }
}

}  // namespace viwb

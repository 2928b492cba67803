// kernels_marg.cuh -- gauge re-anchoring and marginalisation, one block per window.
//
//   reanchor : Estimator::double2vector + vector2double (estimator.cpp:1224-1332, 1155-1222)
//   marg     : MarginalizationInfo::marginalize + getParameterBlocks (marginalization_factor.cpp:183-334) on the
//              dense system assembled by assemble_block(MODE_MARG) from the same Jacobian kernels the solver uses.
// The dropped landmark block is diagonal, so it is eliminated exactly (Schur sum T0 = sum w w^T / a); the remaining
// dropped block (pose 0 + speed-bias 0, or pose 9) goes through the reference's eigen pseudo-inverse (eps 1e-8),
// and the kept n x n system through a parallel cyclic Jacobi eigen-decomposition in shared memory
// (stands in for Eigen::SelfAdjointEigenSolver) to produce J_lin = sqrt(S) V^T, r_lin = sqrt(S^-1) V^T b.
#pragma once
#include "kernels_lin.cuh"
#include "kernels_solve.cuh"

namespace viwb {

VIWB_D void reanchor_block(const BatchDev &bd, int bx, int by, int tid, int nt, double *smem, int mode) {
    (void)by; (void)smem; (void)mode; (void)nt;
    if (tid != 0) return;
    const WinMeta &m = bd.meta[bx];
    double *st = bd.x_cur + m.state_off;
    const double *before = bd.x_before + m.state_off;
    const bool use_imu = (m.flags[BLK_SB0] & 1u) != 0;
    const int nfr = m.frame_count + 1;
    if (use_imu) {
        const M3 Rs0 = qR(ldq(before + 3)), R00 = qR(ldq(st + 3));
        const V3 o0 = R_to_ypr(Rs0), o00 = R_to_ypr(R00), P0 = ld3(before), p00 = ld3(st);
        M3 rot = yaw_to_R(o0.x - o00.x);
        if (fabs(fabs(o0.y) - 90) < 1.0 || fabs(fabs(o00.y) - 90) < 1.0) rot = Rs0 * transpose(R00);
        for (int i = 0; i < nfr; i++) {
            double *p = st + 7 * i;
            const M3 Ri = rot * qR(qnormalized(ldq(p + 3)));
            const V3 Pi = rot * (ld3(p) - p00) + P0;
            st3(p, Pi); stq(p + 3, q_from_R(Ri));
            st3(st + 77 + 9 * i, rot * ld3(st + 77 + 9 * i));
        }
        for (int c = 0; c < 2; c++) if (m.flags[BLK_EX0 + c] & 1u) { double *e = st + blk_off(BLK_EX0 + c); stq(e + 3, q_from_R(qR(qnormalized(ldq(e + 3))))); }
    } else {
        for (int i = 0; i < nfr; i++) { double *p = st + 7 * i; stq(p + 3, q_from_R(qR(qnormalized(ldq(p + 3))))); }
    }
    if (m.flags[BLK_EXW] & 1u) {
        double *e = st + blk_off(BLK_EXW);
        stq(e + 3, q_from_R(qR(qnormalized(ldq(e + 3)))));
        if (m.flags[BLK_PR] & 1u) for (int k = 0; k < 4; k++) st[blk_off(BLK_PR) + k] = e[3 + k];   // quirk 1 (estimator.cpp:1209-1213)
    }
    for (int k = 0; k < m.nlm; k++) st[SFIX + k] = 1.0 / (1.0 / st[SFIX + k]);   // setDepth(1/x), getDepthVector(1/depth)
}

// Symmetric eigen-decomposition in shared memory: Householder tridiagonalisation + implicit QL with eigenvector
// accumulation (the EISPACK tred2 / tql2 pair, the algorithm family Eigen::SelfAdjointEigenSolver uses), with the
// O(n^2)-per-step inner loops spread over the block and the O(n) scalar recurrences kept on thread 0.
// V (n x n, leading dimension ld): in = symmetric matrix (lower triangle read), out = eigenvectors in columns.
// d (n): eigenvalues (unsorted).  e (n), cs (2n), sc (8): scratch.
VIWB_D double blk_reduce_small(double v, int tid, int nt, double *red) {     // sum over the block via a short tree
    return block_sum(v, tid, nt, red);
}
VIWB_D void sym_eig_block(double *V, double *d, double *e, double *cs, double *sc, double *red, int n, int ld, int tid, int nt) {
#define VV(i, j) V[(i) * ld + (j)]
    // ---- tred2
    for (int j = tid; j < n; j += nt) d[j] = VV(n - 1, j);
    VIWB_SYNC();
    for (int i = n - 1; i > 0; i--) {
        double part = 0.0;
        for (int k = tid; k < i; k += nt) part += fabs(d[k]);
        const double scale = blk_reduce_small(part, tid, nt, red);
        if (scale == 0.0) {
            if (tid == 0) e[i] = d[i - 1];
            VIWB_SYNC();
            for (int j = tid; j < i; j += nt) { d[j] = VV(i - 1, j); VV(i, j) = 0.0; VV(j, i) = 0.0; }
            VIWB_SYNC();
            if (tid == 0) d[i] = 0.0;
            VIWB_SYNC();
            continue;
        }
        part = 0.0;
        for (int k = tid; k < i; k += nt) { const double t = d[k] / scale; d[k] = t; part += t * t; }
        double h = blk_reduce_small(part, tid, nt, red);
        if (tid == 0) {
            const double f = d[i - 1];
            double g = sqrt(h); if (f > 0) g = -g;
            e[i] = scale * g; h = h - f * g; d[i - 1] = f - g; sc[0] = h;
        }
        VIWB_SYNC();
        h = sc[0];
        // e[j] = (A d)[j] over the leading i x i block (lower triangle storage); column i keeps the Householder vector
        for (int j = tid; j < i; j += nt) {
            double g = 0.0;
            for (int k = 0; k <= j; k++) g += VV(j, k) * d[k];
            for (int k = j + 1; k < i; k++) g += VV(k, j) * d[k];
            cs[j] = g / h;            // e[j] / h, kept in cs until the reduction below is done
            VV(j, i) = d[j];
        }
        VIWB_SYNC();
        part = 0.0;
        for (int j = tid; j < i; j += nt) part += cs[j] * d[j];
        const double f2 = blk_reduce_small(part, tid, nt, red);
        const double hh = f2 / (h + h);
        for (int j = tid; j < i; j += nt) e[j] = cs[j] - hh * d[j];
        VIWB_SYNC();
        // rank-2 update of the lower triangle: V[k][j] -= d[j] e[k] + e[j] d[k],  j <= k < i
        for (int k = tid; k < i; k += nt) { const double ek = e[k], dk = d[k]; for (int j = 0; j <= k; j++) VV(k, j) -= d[j] * ek + e[j] * dk; }
        VIWB_SYNC();
        for (int j = tid; j < i; j += nt) { cs[j] = VV(i - 1, j); VV(i, j) = 0.0; }
        VIWB_SYNC();
        for (int j = tid; j < i; j += nt) d[j] = cs[j];
        if (tid == 0) d[i] = h;
        VIWB_SYNC();
    }
    // ---- accumulate the transformations
    for (int i = 0; i < n - 1; i++) {
        if (tid == 0) { VV(n - 1, i) = VV(i, i); VV(i, i) = 1.0; }
        VIWB_SYNC();
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = tid; k <= i; k += nt) d[k] = VV(k, i + 1) / h;
            VIWB_SYNC();
            for (int j = tid; j <= i; j += nt) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += VV(k, i + 1) * VV(k, j);
                for (int k = 0; k <= i; k++) VV(k, j) -= g * d[k];
            }
            VIWB_SYNC();
        }
        for (int k = tid; k <= i; k += nt) VV(k, i + 1) = 0.0;
        VIWB_SYNC();
    }
    for (int j = tid; j < n; j += nt) { d[j] = VV(n - 1, j); VV(n - 1, j) = 0.0; }
    VIWB_SYNC();
    if (tid == 0) { VV(n - 1, n - 1) = 1.0; e[0] = 0.0; }
    VIWB_SYNC();
    // ---- tql2
    for (int i = 1 + tid; i < n; i += nt) cs[i - 1] = e[i];
    VIWB_SYNC();
    for (int i = tid; i < n - 1; i += nt) e[i] = cs[i];
    if (tid == 0) { e[n - 1] = 0.0; sc[1] = 0.0 /* f */; sc[2] = 0.0 /* tst1 */; }
    VIWB_SYNC();
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        if (tid == 0) {
            const double t = fabs(d[l]) + fabs(e[l]); if (t > sc[2]) sc[2] = t;
            int m = l; while (m < n) { if (fabs(e[m]) <= eps * sc[2]) break; m++; }
            sc[3] = (double)m;
        }
        VIWB_SYNC();
        const int m = (int)sc[3];
        if (m > l) {
            for (int iter = 0; iter < 200; iter++) {
                if (tid == 0) {
                    double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = sqrt(p * p + 1.0);
                    if (p < 0) r = -r;
                    d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
                    const double dl1 = d[l + 1]; double h = g - d[l];
                    for (int i = l + 2; i < n; i++) d[i] -= h;
                    sc[1] += h;
                    p = d[m];
                    double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0; const double el1 = e[l + 1];
                    // the rotation chain is the serial critical path of the whole decomposition: operands of the next link are
                    // fetched before the current link's sqrt / reciprocal, and one reciprocal replaces two divisions
                    double ei = e[m - 1], di = d[m - 1];
                    for (int i = m - 1; i >= l; i--) {
                        const double ein = i > l ? e[i - 1] : 0.0, din = i > l ? d[i - 1] : 0.0;
                        c3 = c2; c2 = c; s2 = s;
                        g = c * ei; h = c * p;
                        const double q2 = p * p + ei * ei, rinv = q2 > 0.0 ? rsqrt(q2) : 0.0;      // r = q2 * rsqrt(q2): one special-function chain, no division
                        r = q2 * rinv;
                        e[i + 1] = s * r; s = ei * rinv; c = p * rinv;
                        p = c * di - s * g; d[i + 1] = h + s * (c * g + s * di);
                        cs[2 * i] = c; cs[2 * i + 1] = s;
                        ei = ein; di = din;
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p; d[l] = c * p;
                    sc[4] = (fabs(e[l]) > eps * sc[2]) ? 1.0 : 0.0;
                }
                VIWB_SYNC();
                for (int k = tid; k < n; k += nt)
                    for (int i = m - 1; i >= l; i--) { const double c = cs[2 * i], s = cs[2 * i + 1], h = VV(k, i + 1); VV(k, i + 1) = s * VV(k, i) + c * h; VV(k, i) = c * VV(k, i) - s * h; }
                VIWB_SYNC();
                if (sc[4] == 0.0) break;
                VIWB_SYNC();
            }
        }
        if (tid == 0) { d[l] = d[l] + sc[1]; e[l] = 0.0; }
        VIWB_SYNC();
    }
#undef VV
}

VIWB_HD int vsub_to_mlay(int p) { return p < 66 ? p : p < 72 ? 165 + (p - 66) : p < 78 ? 171 + (p - 72) : 191; }   // td -> blk_moff(BLK_TD) = 191
VIWB_HD int marg_cap(int nmax) { return nmax < 16 ? 16 : (nmax > 100 ? 100 : nmax); }
// eigenvector matrix c x (c|1), dropped-block matrix and its pseudo-inverse (<= 16 x 16 each), Arm*Ainv (c x 15), five vectors of <= 2c,
// reduction scratch, scalars, two index lists: 73 KB for the 82-dimensional prior of the stereo+IMU configuration -> three blocks per SM
VIWB_HD size_t marg_smem_doubles(int nt, int nmax) { (void)nt; const int c = marg_cap(nmax); return (size_t)c * (c | 1) + 2 * 256 + (size_t)c * 15 + (size_t)5 * c + 32 + 16 + 108 + 8; }

VIWB_D void marg_block(const BatchDev &bd, int bx, int by, int tid, int nt, double *smem, int mode) {
    (void)by; (void)mode;
    const int w = bx;
    const WinMeta &m = bd.meta[w];
    WinWork &ww = bd.work[w];
    if (m.margin_flag < 0) return;
    const int LDM = MAXPRI + 16;
    double *M = bd.marg_A + (size_t)w * LDM * LDM;
    double *b = bd.gfix + (size_t)w * (TFIX + 8);
    const double *T = bd.Tvis + (size_t)w * VSUB * VSUB, *tv = bd.tvec + (size_t)w * VSUB;
    int *hdr = bd.marg_hdr + (size_t)w * (3 + 2 * NB);
    const double eps = 1e-8;   // marginalization_factor.h:81
    // smem carve
    const int cap = marg_cap(bd.marg_nmax);
    double *Vn = smem, *Amm = Vn + (size_t)cap * (cap | 1), *Ainv = Amm + 256, *Tm = Ainv + 256;
    double *bn = Tm + (size_t)cap * 15, *cs = bn + cap, *ev = cs + 2 * cap, *ee = ev + cap, *red = ee + cap, *bc = red + 32;
    int *keep = (int *)(bc + 16), *dl = keep + 216;
    double *An = Vn;
    // ---- dense system of the marginalisation factors over the marginalisation layout
    for (int e = tid; e < MLAY * MLAY; e += nt) M[(size_t)(e / MLAY) * LDM + (e % MLAY)] = 0.0;
    for (int i = tid; i < TFIX + 8; i += nt) b[i] = 0.0;
    VIWB_SYNC();
    { DenseTarget t; t.M = M; t.g = b; t.ld = LDM; t.flags = m.flags; assemble_into(t, bd, w, MODE_MARG, tid, nt, keep); }      // keep[] is filled only afterwards
    VIWB_SYNC();
    // ---- eliminate the dropped landmarks: M -= scatter(T0), b -= scatter(tvec0)   (MARGIN_OLD only)
    if (m.margin_flag == 0) {
        for (int e = tid; e < 79 * 79; e += nt) { const int p = e / 79, q = e % 79; M[(size_t)vsub_to_mlay(p) * LDM + vsub_to_mlay(q)] -= T[p * VSUB + q]; }
        for (int p = tid; p < 79; p += nt) b[vsub_to_mlay(p)] -= tv[p];
    }
    VIWB_SYNC();
    // ---- dropped / kept dimension lists (marginalisation layout)
    if (tid == 0) {
        int md = 0, n = 0, nb = 0;
        unsigned dropped = 0;
        for (int k = 0; k < SFIX; k++) bd.marg_x0[(size_t)w * SFIX + k] = 0.0;
        if (m.margin_flag == 0) { for (int k = 0; k < 6; k++) dl[md++] = k; for (int k = 0; k < 9; k++) dl[md++] = 66 + k; dropped = (1u << 0) | (1u << BLK_SB0); }
        else { for (int k = 0; k < 6; k++) dl[md++] = 54 + k; dropped = (1u << 9); }
        // a dropped block that no factor references contributes nothing (its rows are zero); keep md as is
        for (int bq = 0; bq < NB; bq++) {
            if (!(m.flags[bq] & 4u) || ((dropped >> bq) & 1u)) continue;      // bit 2 of flags = seen by a marginalisation factor
            int nid = bq;
            if (m.margin_flag == 0) { if ((bq >= 1 && bq <= 10) || (bq >= 12 && bq <= 21)) nid = bq - 1; }
            else { if (bq == 10 || bq == 21) nid = bq - 1; }
            hdr[3 + nb] = nid; hdr[3 + NB + nb] = n; nb++;
            for (int k = 0; k < blk_msize(bq); k++) keep[n++] = blk_moff(bq) + k;
            // linearisation point of the kept block, stored under its new id (keep_block_data + addr_shift)
            const double *src = bd.x_cur + m.state_off + blk_off(bq);
            double *dst = bd.marg_x0 + (size_t)w * SFIX + blk_off(nid);
            for (int k = 0; k < blk_size(bq); k++) dst[k] = src[k];
        }
        hdr[0] = 1; hdr[1] = n; hdr[2] = nb;
        bc[2] = (double)md; bc[3] = (double)n;
    }
    VIWB_SYNC();
    const int md = (int)bc[2], n = (int)bc[3];
    if (n > cap) { if (tid == 0) { ww.marg_status = -1; hdr[0] = 0; } return; }
    // ---- pseudo-inverse of the dropped fixed block (marginalization_factor.cpp:282-287)
    for (int e = tid; e < md * md; e += nt) { const int i = e / md, j = e % md; Amm[e] = 0.5 * (M[(size_t)dl[i] * LDM + dl[j]] + M[(size_t)dl[j] * LDM + dl[i]]); }
    VIWB_SYNC();
    sym_eig_block(Amm, ev, ee, cs, bc + 4, red, md, md, tid, nt);      // eigenvalues -> ev, eigenvectors -> columns of Amm
    VIWB_SYNC();
    for (int e = tid; e < md * md; e += nt) {
        const int i = e / md, j = e % md;
        double sacc = 0.0;
        for (int k = 0; k < md; k++) { const double l = ev[k]; if (l > eps) sacc += Amm[i * md + k] * Amm[j * md + k] / l; }
        Ainv[e] = sacc;
    }
    VIWB_SYNC();
    // Tm = Arm * Ainv (n x md)
    for (int e = tid; e < n * md; e += nt) {
        const int i = e / md, j = e % md;
        double sacc = 0.0;
        for (int k = 0; k < md; k++) sacc += M[(size_t)keep[i] * LDM + dl[k]] * Ainv[k * md + j];
        Tm[e] = sacc;
    }
    VIWB_SYNC();
    // A = Arr - Arm Amm^-1 Amr ; b = brr - Arm Amm^-1 bmm
    const int ld = n | 1;
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        double sacc = M[(size_t)keep[i] * LDM + keep[j]];
        for (int k = 0; k < md; k++) sacc -= Tm[i * md + k] * M[(size_t)dl[k] * LDM + keep[j]];
        An[i * ld + j] = sacc;
    }
    for (int i = tid; i < n; i += nt) {
        double sacc = b[keep[i]];
        for (int k = 0; k < md; k++) sacc -= Tm[i * md + k] * b[dl[k]];
        bn[i] = sacc;
    }
    VIWB_SYNC();
    // SelfAdjointEigenSolver reads the lower triangle
    for (int e = tid; e < n * n; e += nt) { const int i = e / n, j = e % n; if (j > i) An[i * ld + j] = An[j * ld + i]; }
    VIWB_SYNC();
    sym_eig_block(Vn, ev, ee, cs, bc + 4, red, n, ld, tid, nt);
    VIWB_SYNC();
    // J_lin = sqrt(S) V^T, r_lin = sqrt(S^-1) V^T b   (marginalization_factor.cpp:298-306)
    double *Jout = bd.marg_J + (size_t)w * bd.marg_nmax * bd.marg_nmax, *rout = bd.marg_r + (size_t)w * MAXPRI;
    for (int i = tid; i < n; i += nt) {
        const double l = ev[i];
        const double S = l > eps ? l : 0.0, Sinv = l > eps ? 1.0 / l : 0.0, ss = sqrt(S), si = sqrt(Sinv);
        double vb = 0.0;
        for (int k = 0; k < n; k++) { Jout[(size_t)i * n + k] = ss * Vn[k * ld + i]; vb += Vn[k * ld + i] * bn[k]; }
        rout[i] = si * vb;
    }
    if (tid == 0) ww.marg_status = 0;
}

// ---------------------------------------------------------------------------------------------------- outlier rejection
// Estimator::outliersRejection (estimator.cpp:2127-2185) on the solved window: per landmark the mean of
// reprojectionError (:2115-2125) over its observations -- left camera of another frame, right camera of another frame, right
// camera of the host frame; the raw normalised points, no td compensation -- times FOCAL_LENGTH against 3 px.
// One thread per landmark (its factors are consecutive in the table).
VIWB_D void outlier_block(const BatchDev &bd, int bx, int by, int tid, int nt, double *smem, int mode) {
    (void)by; (void)smem; (void)mode;
    const int k = bx * nt + tid;
    if (k >= bd.nlm_total) return;
    const int w = bd.lm_win[k];
    const WinMeta &m = bd.meta[w];
    const double *x = bd.x_cur + m.state_off;
    const int f0 = bd.lm_fptr[k], f1 = bd.lm_fptr[k + 1];
    if (f0 == f1) { bd.lm_outlier[k] = 0; return; }
    const double depth = 1.0 / x[SFIX + (k - m.lm_off)];
    const int ex_off[2] = {blk_off(BLK_EX0), blk_off(BLK_EX1)};
    double err = 0.0; int cnt = 0;
    for (int f = f0; f < f1; f++) {
        const int type = bd.vis_type[f], i = bd.vis_fi[f], j = (type == 2) ? i : bd.vis_fj[f], cam = (type == 0) ? 0 : 1;
        const double *o = bd.vis_obs + (size_t)f * 12;
        const V3 uvi = ld3(o), uvj = ld3(o + 3);
        const V3 Pi = ld3(x + 7 * i), Pj = ld3(x + 7 * j), tici = ld3(x + ex_off[0]), ticj = ld3(x + ex_off[cam]);
        const Q4 Qi = ldq(x + 7 * i + 3), Qj = ldq(x + 7 * j + 3), qici = ldq(x + ex_off[0] + 3), qicj = ldq(x + ex_off[cam] + 3);
        const V3 pts_w = qrot(Qi, qrot(qici, depth * uvi) + tici) + Pi;
        const V3 pts_cj = tmul(qR(qicj), tmul(qR(Qj), pts_w - Pj) - ticj);
        const double rx = pts_cj.x / pts_cj.z - uvj.x, ry = pts_cj.y / pts_cj.z - uvj.y;
        err += sqrt(rx * rx + ry * ry); cnt++;
    }
    bd.lm_outlier[k] = ((err / cnt) * bd.out_focal > bd.out_thresh) ? 1 : 0;
}

}  // namespace viwb

/*
 * vo_marg.c -- CPU oracle, part 3: marginalization, gauge re-anchoring, Estimator::optimization().
 * TEST INFRASTRUCTURE ONLY (see viw_oracle.h).
 * Restates factor/marginalization_factor.cpp:12-334 and estimator/estimator.cpp:1155-1332,1669-1893.
 * Eigen::SelfAdjointEigenSolver (third party) is replaced by Householder tridiagonalisation + implicit QL
 * (the same algorithm family; EISPACK tred2/tql2).
 */
#include "viw_oracle.h"
#include "vo_math.h"
#include <stdlib.h>

/* ------------------------------------------------------------------ symmetric eigen solver */
int vo_sym_eig(int n, const double *A, double *d, double *Vout) {
    if (n <= 0) return 0;
    double *V = Vout, *e = (double *)malloc(sizeof(double) * n);
    /* SelfAdjointEigenSolver reads the lower triangle only */
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (j <= i) ? A[i * n + j] : A[j * n + i];
#define VV(i, j) V[(i) * n + (j)]
    for (int j = 0; j < n; j++) d[j] = VV(n - 1, j);
    for (int i = n - 1; i > 0; i--) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; k++) scale += fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; j++) { d[j] = VV(i - 1, j); VV(i, j) = 0.0; VV(j, i) = 0.0; }
        } else {
            for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1], g = sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g; h = h - f * g; d[i - 1] = f - g;
            for (int j = 0; j < i; j++) e[j] = 0.0;
            for (int j = 0; j < i; j++) {
                f = d[j]; VV(j, i) = f; g = e[j] + VV(j, j) * f;
                for (int k = j + 1; k <= i - 1; k++) { g += VV(k, j) * d[k]; e[k] += VV(k, j) * f; }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
            double hh = f / (h + h);
            for (int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (int j = 0; j < i; j++) {
                f = d[j]; g = e[j];
                for (int k = j; k <= i - 1; k++) VV(k, j) -= (f * e[k] + g * d[k]);
                d[j] = VV(i - 1, j); VV(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; i++) {
        VV(n - 1, i) = VV(i, i); VV(i, i) = 1.0;
        double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; k++) d[k] = VV(k, i + 1) / h;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += VV(k, i + 1) * VV(k, j);
                for (int k = 0; k <= i; k++) VV(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; k++) VV(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; j++) { d[j] = VV(n - 1, j); VV(n - 1, j) = 0.0; }
    VV(n - 1, n - 1) = 1.0; e[0] = 0.0;
    /* tql2 */
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0; const double eps = 2.220446049250313e-16;
    int rc = 0;
    for (int l = 0; l < n; l++) {
        double t = fabs(d[l]) + fabs(e[l]); if (t > tst1) tst1 = t;
        int m = l;
        while (m < n) { if (fabs(e[m]) <= eps * tst1) break; m++; }
        if (m > l) {
            int iter = 0;
            do {
                if (++iter > 200) { rc = -1; break; }
                double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
                double dl1 = d[l + 1], h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i]; h = c * p; r = hypot(p, e[i]);
                    e[i + 1] = s * r; s = e[i] / r; c = p / r;
                    p = c * d[i] - s * g; d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) { h = VV(k, i + 1); VV(k, i + 1) = s * VV(k, i) + c * h; VV(k, i) = c * VV(k, i) - s * h; }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p; d[l] = c * p;
            } while (fabs(e[l]) > eps * tst1);
        }
        d[l] = d[l] + f; e[l] = 0.0;
    }
    /* sort ascending (SelfAdjointEigenSolver returns increasing eigenvalues) */
    for (int i = 0; i < n - 1; i++) {
        int k = i; double p = d[i];
        for (int j = i + 1; j < n; j++) if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) { d[k] = d[i]; d[i] = p; for (int j = 0; j < n; j++) { double t = VV(j, i); VV(j, i) = VV(j, k); VV(j, k) = t; } }
    }
#undef VV
    free(e);
    return rc;
}

/* ------------------------------------------------------------------ gauge re-anchoring */
/* Estimator::double2vector followed by vector2double (estimator.cpp:1224-1332, 1155-1222). */
int vo_gauge_reanchor(const viwb_problem *pb, const double *before, double *st) {
    int use_imu = (pb->block_flags[VIWB_BLK_SPEEDBIAS0] & VIWB_BLOCK_PRESENT) != 0;
    int nfr = pb->frame_count + 1;
    double Rs[VIWB_NUM_FRAMES][9], Ps[VIWB_NUM_FRAMES][3], Vs[VIWB_NUM_FRAMES][3];
    if (use_imu) {
        double Rs0[9], origin_R0[3], origin_P0[3], R00[9], origin_R00[3], rot_diff[9], q[4];
        q_to_R(Rs0, before + 3); R_to_ypr(origin_R0, Rs0); v3_copy(origin_P0, before);
        q_to_R(R00, st + 3); R_to_ypr(origin_R00, R00);
        double y_diff = origin_R0[0] - origin_R00[0];
        double ypr[3] = {y_diff, 0, 0}; ypr_to_R(rot_diff, ypr);
        if (fabs(fabs(origin_R0[1]) - 90) < 1.0 || fabs(fabs(origin_R00[1]) - 90) < 1.0) {
            double t[9]; m3_transpose(t, R00); m3_mul(rot_diff, Rs0, t);
        }
        for (int i = 0; i < nfr; i++) {
            const double *p = st + 7 * i; double Ri[9], d[3];
            memcpy(q, p + 3, sizeof q); q_normalize(q); q_to_R(Ri, q); m3_mul(Rs[i], rot_diff, Ri);
            d[0] = p[0] - st[0]; d[1] = p[1] - st[1]; d[2] = p[2] - st[2];
            m3_mulv(Ps[i], rot_diff, d); v3_add(Ps[i], Ps[i], origin_P0);
            m3_mulv(Vs[i], rot_diff, st + 77 + 9 * i);
        }
    } else {
        for (int i = 0; i < nfr; i++) { double q[4]; memcpy(q, st + 7 * i + 3, sizeof q); q_normalize(q); q_to_R(Rs[i], q); v3_copy(Ps[i], st + 7 * i); }
    }
    /* vector2double */
    for (int i = 0; i < nfr; i++) {
        v3_copy(st + 7 * i, Ps[i]); q_from_R(st + 7 * i + 3, Rs[i]);
        if (use_imu) v3_copy(st + 77 + 9 * i, Vs[i]);
    }
    if (use_imu) for (int c = 0; c < 2; c++) if (pb->block_flags[VIWB_BLK_EX_POSE0 + c] & VIWB_BLOCK_PRESENT) {
        double *e = st + viwb_block_offset(VIWB_BLK_EX_POSE0 + c), q[4], R[9];
        memcpy(q, e + 3, sizeof q); q_normalize(q); q_to_R(R, q); q_from_R(e + 3, R);
    }
    if (pb->block_flags[VIWB_BLK_EX_WHEEL] & VIWB_BLOCK_PRESENT) {
        double *e = st + viwb_block_offset(VIWB_BLK_EX_WHEEL), q[4], R[9];
        memcpy(q, e + 3, sizeof q); q_normalize(q); q_to_R(R, q); q_from_R(e + 3, R);
        /* quirk 1 (estimator.cpp:1209-1213): para_plane_R is refilled from the wheel extrinsic quaternion */
        if (pb->block_flags[VIWB_BLK_PLANE_R] & VIWB_BLOCK_PRESENT) memcpy(st + viwb_block_offset(VIWB_BLK_PLANE_R), e + 3, 4 * sizeof(double));
    }
    /* setDepth(1/x) then getDepthVector(1/depth) (feature_manager.cpp:142-160,179-195) */
    for (int k = 0; k < pb->num_landmarks; k++) st[VIWB_STATE_FIXED + k] = 1.0 / (1.0 / st[VIWB_STATE_FIXED + k]);
    return 0;
}

/* ------------------------------------------------------------------ marginalization */
typedef struct { int type, nrows, nslots, block[32]; const double *consts; int has_loss; int drop[32]; } minfo_t;

static int marg_local(int b) { return b < VIWB_NUM_FIXED_BLOCKS ? viwb_block_marg_size(b) : 1; }
static int gsize_of(int b) { return b < VIWB_NUM_FIXED_BLOCKS ? viwb_block_size(b) : 1; }
static int soff_of(int b) { return b < VIWB_NUM_FIXED_BLOCKS ? viwb_block_offset(b) : VIWB_STATE_FIXED + (b - VIWB_NUM_FIXED_BLOCKS); }

int vo_marginalize(const viwb_problem *pb, const double *state, int flag, viwb_prior *out,
                   double *A_out, double *b_out, int32_t *mn_out) {
    const int has_prior = pb->prior && pb->prior->valid;
    const int NB = VIWB_NUM_FIXED_BLOCKS + pb->num_landmarks;
    int nf_max = 1 + 3 + pb->num_vis;
    minfo_t *F = (minfo_t *)calloc(nf_max, sizeof(minfo_t));
    int nfac = 0;
    if (flag == VIWB_MARGIN_OLD) {
        if (has_prior) {
            minfo_t *f = &F[nfac++]; f->type = -1; f->nrows = pb->prior->n; f->nslots = pb->prior->num_blocks;
            for (int i = 0; i < f->nslots; i++) { f->block[i] = pb->prior->block_id[i]; f->drop[i] = (f->block[i] == VIWB_BLK_POSE0 || f->block[i] == VIWB_BLK_SPEEDBIAS0); }
        }
        for (int i = 0; i < pb->num_imu; i++) if (pb->imu_frame_i[i] == 0 && pb->imu_frame_j[i] == 1) {
            minfo_t *f = &F[nfac++]; f->type = VIWB_F_IMU; f->nrows = 15; f->nslots = 4; f->consts = pb->imu_data + (size_t)i * VIWB_IMU_DOUBLES;
            int b[4] = {VIWB_BLK_POSE0, VIWB_BLK_SPEEDBIAS0, VIWB_BLK_POSE0 + 1, VIWB_BLK_SPEEDBIAS0 + 1};
            for (int k = 0; k < 4; k++) f->block[k] = b[k];
            f->drop[0] = f->drop[1] = 1;
        }
        for (int i = 0; i < pb->num_wheel; i++) if (pb->wheel_frame_i[i] == 0 && pb->wheel_frame_j[i] == 1) {
            minfo_t *f = &F[nfac++]; f->type = VIWB_F_WHEEL; f->nrows = 6; f->nslots = 7; f->consts = pb->wheel_data + (size_t)i * VIWB_WHEEL_DOUBLES;
            int b[7] = {VIWB_BLK_POSE0, VIWB_BLK_POSE0 + 1, VIWB_BLK_EX_WHEEL, VIWB_BLK_SX, VIWB_BLK_SY, VIWB_BLK_SW, VIWB_BLK_TD_WHEEL};
            for (int k = 0; k < 7; k++) f->block[k] = b[k];
            f->drop[0] = 1;
        }
        for (int i = 0; i < pb->num_plane; i++) if (pb->plane_frame[i] == 0) {
            minfo_t *f = &F[nfac++]; f->type = VIWB_F_PLANE; f->nrows = 3; f->nslots = 4;
            int b[4] = {VIWB_BLK_POSE0, VIWB_BLK_EX_WHEEL, VIWB_BLK_PLANE_R, VIWB_BLK_PLANE_Z};
            for (int k = 0; k < 4; k++) f->block[k] = b[k];
            f->drop[0] = 1;
        }
        for (int i = 0; i < pb->num_vis; i++) {
            if (pb->vis_frame_i[i] != 0) continue;
            int fi = 0, fj = pb->vis_frame_j[i], lm = VIWB_BLK_LANDMARK0 + pb->vis_landmark[i];
            minfo_t *f = &F[nfac++]; f->type = pb->vis_type[i]; f->nrows = 2; f->has_loss = 1; f->consts = pb->vis_obs + (size_t)i * VIWB_VIS_OBS_DOUBLES;
            if (f->type == VIWB_F_PROJ_2F1C) { int b[5] = {VIWB_BLK_POSE0 + fi, VIWB_BLK_POSE0 + fj, VIWB_BLK_EX_POSE0, lm, VIWB_BLK_TD}; f->nslots = 5; for (int k = 0; k < 5; k++) f->block[k] = b[k]; f->drop[0] = 1; f->drop[3] = 1; }
            else if (f->type == VIWB_F_PROJ_2F2C) { int b[6] = {VIWB_BLK_POSE0 + fi, VIWB_BLK_POSE0 + fj, VIWB_BLK_EX_POSE0, VIWB_BLK_EX_POSE1, lm, VIWB_BLK_TD}; f->nslots = 6; for (int k = 0; k < 6; k++) f->block[k] = b[k]; f->drop[0] = 1; f->drop[4] = 1; }
            else { int b[4] = {VIWB_BLK_EX_POSE0, VIWB_BLK_EX_POSE1, lm, VIWB_BLK_TD}; f->nslots = 4; for (int k = 0; k < 4; k++) f->block[k] = b[k]; f->drop[2] = 1; }
        }
    } else {
        int has9 = 0;
        if (has_prior) for (int i = 0; i < pb->prior->num_blocks; i++) if (pb->prior->block_id[i] == VIWB_BLK_POSE0 + VIWB_WINDOW_SIZE - 1) has9 = 1;
        if (!has9) {
            /* estimator.cpp:1821-1822: prior untouched */
            if (has_prior) {
                const viwb_prior *p = pb->prior; out->valid = 1; out->n = p->n; out->num_blocks = p->num_blocks;
                memcpy(out->block_id, p->block_id, sizeof p->block_id); memcpy(out->block_idx, p->block_idx, sizeof p->block_idx);
                memcpy(out->x0, p->x0, sizeof(double) * VIWB_STATE_FIXED); memcpy(out->J, p->J, sizeof(double) * p->n * p->n); memcpy(out->r, p->r, sizeof(double) * p->n);
            } else { out->valid = 0; out->n = 0; out->num_blocks = 0; }
            if (mn_out) { mn_out[0] = 0; mn_out[1] = out->n; }
            free(F); return 0;
        }
        minfo_t *f = &F[nfac++]; f->type = -1; f->nrows = pb->prior->n; f->nslots = pb->prior->num_blocks;
        for (int i = 0; i < f->nslots; i++) { f->block[i] = pb->prior->block_id[i]; f->drop[i] = (f->block[i] == VIWB_BLK_POSE0 + VIWB_WINDOW_SIZE - 1); }
    }
    /* addResidualBlockInfo: parameter_block_size / parameter_block_idx (marginalization_factor.cpp:98-117) */
    int *seen = (int *)calloc(NB, sizeof(int)), *dropped = (int *)calloc(NB, sizeof(int)), *idx = (int *)malloc(sizeof(int) * NB);
    for (int i = 0; i < nfac; i++) for (int s = 0; s < F[i].nslots; s++) { seen[F[i].block[s]] = 1; if (F[i].drop[s]) dropped[F[i].block[s]] = 1; }
    int pos = 0;
    for (int b = 0; b < NB; b++) if (seen[b] && dropped[b]) { idx[b] = pos; pos += marg_local(b); }
    int m = pos;
    for (int b = 0; b < NB; b++) if (seen[b] && !dropped[b]) { idx[b] = pos; pos += marg_local(b); }
    int n = pos - m;
    if (mn_out) { mn_out[0] = m; mn_out[1] = n; }
    if (m == 0) { out->valid = 0; out->n = 0; out->num_blocks = 0; free(F); free(seen); free(dropped); free(idx); return 0; }
    if (n > VIWB_MAX_PRIOR_DIM) { free(F); free(seen); free(dropped); free(idx); return VIWB_ERR_INVALID; }
    /* preMarginalize + ThreadsConstructA (:119-181): A += Ji^T Jj, b += Ji^T r */
    double *A = (double *)calloc((size_t)pos * pos, sizeof(double)), *bv = (double *)calloc(pos, sizeof(double));
    double *pj = NULL; double gj[7][135], res_small[15];
    for (int i = 0; i < nfac; i++) {
        minfo_t *f = &F[i];
        double *res; const double *Jslot[32]; int ld[32];
        double *res_big = NULL;
        if (f->type < 0) {
            if (!pj) pj = (double *)malloc(sizeof(double) * pb->prior->n * VIWB_STATE_FIXED);
            res_big = (double *)malloc(sizeof(double) * f->nrows); res = res_big;
            vo_prior_evaluate(pb->prior, state, res, pj);
            for (int s = 0; s < f->nslots; s++) { Jslot[s] = pj + viwb_block_offset(f->block[s]); ld[s] = VIWB_STATE_FIXED; }
        } else {
            const double *params[7]; double *jp[7];
            for (int s = 0; s < f->nslots; s++) { params[s] = state + soff_of(f->block[s]); jp[s] = gj[s]; }
            res = res_small;
            int rc = vo_factor_evaluate(f->type, &pb->globals, f->consts, params, res, jp);
            if (rc) { free(A); free(bv); free(F); free(seen); free(dropped); free(idx); free(pj); return rc; }
            for (int s = 0; s < f->nslots; s++) { Jslot[s] = gj[s]; ld[s] = gsize_of(f->block[s]); }
            if (f->has_loss) { /* ResidualBlockInfo::Evaluate loss correction (:46-77) on the global Jacobians */
                double sq = 0; for (int r = 0; r < f->nrows; r++) sq += res[r] * res[r];
                double rho[3]; vo_huber(pb->globals.huber_delta, sq, rho);
                double sqrt_rho1 = sqrt(rho[1]), residual_scaling, alpha_sq_norm;
                if (sq == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
                else { double D = 1.0 + 2.0 * sq * rho[2] / rho[1], alpha = 1.0 - sqrt(D); residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / sq; }
                for (int s = 0; s < f->nslots; s++) { int gs = ld[s];
                    for (int c = 0; c < gs; c++) {
                        double rtj = 0; for (int r = 0; r < f->nrows; r++) rtj += res[r] * gj[s][r * gs + c];
                        for (int r = 0; r < f->nrows; r++) gj[s][r * gs + c] = sqrt_rho1 * (gj[s][r * gs + c] - alpha_sq_norm * res[r] * rtj);
                    } }
                for (int r = 0; r < f->nrows; r++) res[r] *= residual_scaling;
            }
        }
        for (int a = 0; a < f->nslots; a++) {
            int ia = idx[f->block[a]], sa = marg_local(f->block[a]);
            for (int c = 0; c < sa; c++) { double v = 0; for (int r = 0; r < f->nrows; r++) v += Jslot[a][r * ld[a] + c] * res[r]; bv[ia + c] += v; }
            for (int b2 = a; b2 < f->nslots; b2++) {
                int ib = idx[f->block[b2]], sb = marg_local(f->block[b2]);
                for (int c = 0; c < sa; c++) for (int d = 0; d < sb; d++) {
                    double v = 0; for (int r = 0; r < f->nrows; r++) v += Jslot[a][r * ld[a] + c] * Jslot[b2][r * ld[b2] + d];
                    A[(size_t)(ia + c) * pos + ib + d] += v;
                    if (a != b2) A[(size_t)(ib + d) * pos + ia + c] += v;
                }
            }
        }
        free(res_big);
    }
    free(pj);
    if (A_out) memcpy(A_out, A, sizeof(double) * (size_t)pos * pos);
    if (b_out) memcpy(b_out, bv, sizeof(double) * pos);
    /* marginalize (:282-306) */
    const double eps = 1e-8;  /* marginalization_factor.h:81 */
    double *Amm = (double *)malloc(sizeof(double) * (size_t)m * m * 3), *Vm = Amm + (size_t)m * m, *Ainv = Vm + (size_t)m * m, *wm = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]);
    vo_sym_eig(m, Amm, wm, Vm);
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) {
        double v = 0; for (int k = 0; k < m; k++) if (wm[k] > eps) v += Vm[(size_t)i * m + k] * (1.0 / wm[k]) * Vm[(size_t)j * m + k];
        Ainv[(size_t)i * m + j] = v;
    }
    /* T = Arm * Amm_inv (n x m) */
    double *T = (double *)malloc(sizeof(double) * (size_t)n * m), *An = (double *)malloc(sizeof(double) * (size_t)n * n * 2), *Vn = An + (size_t)n * n;
    double *bn = (double *)malloc(sizeof(double) * n * 2), *wn = bn + n;
    for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) { double v = 0; for (int k = 0; k < m; k++) v += A[(size_t)(m + i) * pos + k] * Ainv[(size_t)k * m + j]; T[(size_t)i * m + j] = v; }
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) { double v = 0; for (int k = 0; k < m; k++) v += T[(size_t)i * m + k] * A[(size_t)k * pos + m + j]; An[(size_t)i * n + j] = A[(size_t)(m + i) * pos + m + j] - v; }
        double v = 0; for (int k = 0; k < m; k++) v += T[(size_t)i * m + k] * bv[k];
        bn[i] = bv[m + i] - v;
    }
    vo_sym_eig(n, An, wn, Vn);
    out->valid = 1; out->n = n;
    for (int i = 0; i < n; i++) {
        double S = wn[i] > eps ? wn[i] : 0, Sinv = wn[i] > eps ? 1.0 / wn[i] : 0, ss = sqrt(S), si = sqrt(Sinv), vb = 0;
        for (int k = 0; k < n; k++) { out->J[(size_t)i * n + k] = ss * Vn[(size_t)k * n + i]; vb += Vn[(size_t)k * n + i] * bn[k]; }
        out->r[i] = si * vb;
    }
    /* getParameterBlocks with addr_shift (:314-334; estimator.cpp:1791-1811, 1866-1888) */
    int nb = 0;
    memset(out->x0, 0, sizeof(double) * VIWB_STATE_FIXED);
    for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) {
        if (!seen[b] || dropped[b]) continue;
        int nb_id = b;
        if (flag == VIWB_MARGIN_OLD) { if (b >= 1 && b <= 10) nb_id = b - 1; else if (b >= 12 && b <= 21) nb_id = b - 1; }
        else { if (b == VIWB_BLK_POSE0 + VIWB_WINDOW_SIZE) nb_id = b - 1; else if (b == VIWB_BLK_SPEEDBIAS0 + VIWB_WINDOW_SIZE) nb_id = b - 1; }
        out->block_id[nb] = nb_id; out->block_idx[nb] = idx[b] - m;
        memcpy(out->x0 + viwb_block_offset(nb_id), state + viwb_block_offset(b), sizeof(double) * viwb_block_size(b));
        nb++;
    }
    out->num_blocks = nb;
    free(Amm); free(wm); free(T); free(An); free(bn); free(A); free(bv); free(F); free(seen); free(dropped); free(idx);
    return 0;
}

int vo_optimization(const viwb_problem *pb, double *state, const viwb_options *opt, int flag, viwb_summary *sum, viwb_prior *out) {
    int n = VIWB_STATE_FIXED + pb->num_landmarks;
    double *before = (double *)malloc(sizeof(double) * n);
    memcpy(before, state, sizeof(double) * n);
    int rc = vo_window_solve(pb, state, opt, sum, NULL);
    if (!rc) rc = vo_gauge_reanchor(pb, before, state);
    free(before);
    if (rc) return rc;
    if (pb->frame_count < VIWB_WINDOW_SIZE || !out) return 0;          /* estimator.cpp:1666 */
    return vo_marginalize(pb, state, flag, out, NULL, NULL, NULL);
}


/* ------------------------------------------------------------------ Estimator::outliersRejection (estimator.cpp:2115-2185) */
static double vo_reproj_err(const double *Pi, const double *Qi, const double *tici, const double *qici, const double *Pj, const double *Qj,
                            const double *ticj, const double *qicj, double depth, const double *uvi, const double *uvj) {
    double t[3], a[3], pw[3], Rj[9], Rc[9], c[3], d[3];
    for (int k = 0; k < 3; k++) t[k] = depth * uvi[k];
    q_rot(a, qici, t); for (int k = 0; k < 3; k++) a[k] += tici[k];
    q_rot(pw, Qi, a); for (int k = 0; k < 3; k++) pw[k] += Pi[k];
    q_to_R(Rj, Qj); q_to_R(Rc, qicj);
    for (int k = 0; k < 3; k++) d[k] = pw[k] - Pj[k];
    for (int r = 0; r < 3; r++) c[r] = Rj[0 * 3 + r] * d[0] + Rj[1 * 3 + r] * d[1] + Rj[2 * 3 + r] * d[2] - ticj[r];     /* Rj^T (pw - Pj) - ticj */
    for (int r = 0; r < 3; r++) a[r] = Rc[0 * 3 + r] * c[0] + Rc[1 * 3 + r] * c[1] + Rc[2 * 3 + r] * c[2];              /* ricj^T (...) */
    const double rx = a[0] / a[2] - uvj[0], ry = a[1] / a[2] - uvj[1];
    return sqrt(rx * rx + ry * ry);
}
int vo_outlier_rejection(const viwb_problem *pb, const double *state, double focal, double thresh, uint8_t *out) {
    const int N = pb->num_landmarks;
    double *err = (double *)calloc(N > 0 ? N : 1, sizeof(double)); int *cnt = (int *)calloc(N > 0 ? N : 1, sizeof(int));
    for (int f = 0; f < pb->num_vis; f++) {
        const int type = pb->vis_type[f], k = pb->vis_landmark[f], i = pb->vis_frame_i[f], j = type == 2 ? i : pb->vis_frame_j[f], cam = type == 0 ? 0 : 1;
        const double *o = pb->vis_obs + (size_t)f * 12, *ex0 = state + 176, *exc = state + 176 + 7 * cam;
        err[k] += vo_reproj_err(state + 7 * i, state + 7 * i + 3, ex0, ex0 + 3, state + 7 * j, state + 7 * j + 3, exc, exc + 3, 1.0 / state[VIWB_STATE_FIXED + k], o, o + 3);
        cnt[k]++;
    }
    for (int k = 0; k < N; k++) out[k] = (cnt[k] > 0 && (err[k] / cnt[k]) * focal > thresh) ? 1 : 0;
    free(err); free(cnt);
    return 0;
}


/* ------------------------------------------------------------------ FeatureManager::triangulate / removeBackShiftDepth
 * triangulatePoint (feature_manager.cpp:198-213) takes the right singular vector of the smallest singular value of the 4x4
 * design matrix from Eigen's JacobiSVD; restated with a one-sided Jacobi SVD (any accurate SVD returns the same vector up to
 * sign, which cancels in the division by its 4th component). */
static void vo_min_right_singular4(double *D, double *v) {
    double V[16];
    for (int i = 0; i < 16; i++) V[i] = (i / 4 == i % 4) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        int rot = 0;
        for (int p = 0; p < 3; p++) for (int q = p + 1; q < 4; q++) {
            double a = 0, b = 0, c = 0;
            for (int r = 0; r < 4; r++) { a += D[r * 4 + p] * D[r * 4 + p]; b += D[r * 4 + q] * D[r * 4 + q]; c += D[r * 4 + p] * D[r * 4 + q]; }
            if (c == 0.0 || fabs(c) <= 1e-300 + 2.220446049250313e-16 * sqrt(a * b)) continue;
            rot++;
            double zeta = (b - a) / (2.0 * c), t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
            for (int r = 0; r < 4; r++) {
                double dp = D[r * 4 + p], dq = D[r * 4 + q]; D[r * 4 + p] = cs * dp - sn * dq; D[r * 4 + q] = sn * dp + cs * dq;
                double vp = V[r * 4 + p], vq = V[r * 4 + q]; V[r * 4 + p] = cs * vp - sn * vq; V[r * 4 + q] = sn * vp + cs * vq;
            }
        }
        if (!rot) break;
    }
    int best = 0; double nb = 0;
    for (int c = 0; c < 4; c++) { double n2 = 0; for (int r = 0; r < 4; r++) n2 += D[r * 4 + c] * D[r * 4 + c]; if (c == 0 || n2 < nb) { nb = n2; best = c; } }
    for (int r = 0; r < 4; r++) v[r] = V[r * 4 + best];
}
static void vo_cam_pose34(const double *pose, const double *ex, double *P) {
    double Rs[9], ric[9], R[9], t[3], a[3];
    q_to_R(Rs, pose + 3); q_to_R(ric, ex + 3); m3_mul(R, Rs, ric);
    m3_mulv(a, Rs, ex); for (int k = 0; k < 3; k++) t[k] = pose[k] + a[k];
    for (int r = 0; r < 3; r++) { double s = 0; for (int c = 0; c < 3; c++) { P[r * 4 + c] = R[c * 3 + r]; s += R[c * 3 + r] * t[c]; } P[r * 4 + 3] = -s; }
}
int vo_triangulate(const double *state, int n, const int32_t *stereo, const int32_t *frame, const double *pt0, const double *pt1, double init_depth, double *depth) {
    for (int k = 0; k < n; k++) {
        double P0[12], P1[12], D[16], v[4];
        const int i = frame[k];
        vo_cam_pose34(state + 7 * i, state + 176, P0);
        if (stereo[k]) vo_cam_pose34(state + 7 * i, state + 183, P1); else vo_cam_pose34(state + 7 * (i + 1), state + 176, P1);
        for (int c = 0; c < 4; c++) {
            D[c] = pt0[2 * k] * P0[8 + c] - P0[c]; D[4 + c] = pt0[2 * k + 1] * P0[8 + c] - P0[4 + c];
            D[8 + c] = pt1[2 * k] * P1[8 + c] - P1[c]; D[12 + c] = pt1[2 * k + 1] * P1[8 + c] - P1[4 + c];
        }
        vo_min_right_singular4(D, v);
        const double px = v[0] / v[3], py = v[1] / v[3], pz = v[2] / v[3];
        const double d = P0[8] * px + P0[9] * py + P0[10] * pz + P0[11];
        depth[k] = d > 0 ? d : init_depth;
    }
    return 0;
}
int vo_shift_depth(int n, const double *uv, const double *depth_in, const double *margR, const double *margP, const double *newR, const double *newP,
                   double init_depth, double *depth_out) {
    for (int k = 0; k < n; k++) {
        double p[3] = {uv[3 * k] * depth_in[k], uv[3 * k + 1] * depth_in[k], uv[3 * k + 2] * depth_in[k]}, w[3], d[3];
        m3_mulv(w, margR, p); for (int c = 0; c < 3; c++) d[c] = w[c] + margP[c] - newP[c];
        const double z = newR[0 * 3 + 2] * d[0] + newR[1 * 3 + 2] * d[1] + newR[2 * 3 + 2] * d[2];      /* (new_R^T d).z */
        depth_out[k] = z > 0 ? z : init_depth;
    }
    return 0;
}

/* ------------------------------------------------------------------ multi-threaded batch (CPU baseline arm of bench.py)
 * n independent windows on `threads` pthreads; each optimisation itself is single-threaded like Ceres' default
 * (estimator.cpp:1646 leaves num_threads commented out).  repeat > 1 cycles over the windows to fill a time budget. */
#include <pthread.h>
#include <malloc.h>
typedef struct { const viwb_problem *pb; const double *const *states; const int32_t *flags; const viwb_options *opt; int n, repeat, tid, nthreads; long done;
                 double *const *out_states; int32_t *out_iters; double *out_prior; } vo_job;
static void *vo_worker(void *arg) {
    vo_job *j = (vo_job *)arg;
    viwb_prior out; double *x0 = (double *)malloc(sizeof(double) * VIWB_STATE_FIXED), *J = (double *)malloc(sizeof(double) * VIWB_MAX_PRIOR_DIM * VIWB_MAX_PRIOR_DIM), *r = (double *)malloc(sizeof(double) * VIWB_MAX_PRIOR_DIM);
    out.x0 = x0; out.J = J; out.r = r;
    double *st = (double *)malloc(sizeof(double) * (VIWB_STATE_FIXED + VIWB_MAX_LANDMARKS));
    viwb_summary sum;
    for (int rep = 0; rep < j->repeat; rep++)
        for (int i = j->tid; i < j->n; i += j->nthreads) {
            memcpy(st, j->states[i], sizeof(double) * (VIWB_STATE_FIXED + j->pb[i].num_landmarks));
            vo_optimization(&j->pb[i], st, j->opt, j->flags ? j->flags[i] : -1, &sum, j->flags ? &out : NULL);
            j->done++;
            /* optional results (bench.py's full-batch parity): solved state, iteration count, and the order-independent content of the new
             * prior: n, trace(J^T J) = |J|_F^2 and |J^T r|^2 (rows of J_lin are eigen-directions: their order and signs are not defined) */
            if (j->out_states && j->out_states[i]) memcpy(j->out_states[i], st, sizeof(double) * (VIWB_STATE_FIXED + j->pb[i].num_landmarks));
            if (j->out_iters) j->out_iters[i] = sum.num_iterations;
            if (j->out_prior) {
                double *o = j->out_prior + 3 * (size_t)i; o[0] = o[1] = o[2] = 0.0;
                if (j->flags && out.valid) {
                    const int n = out.n; double tr = 0.0, g2 = 0.0;
                    for (int a = 0; a < n * n; a++) tr += J[a] * J[a];
                    for (int c = 0; c < n; c++) { double g = 0.0; for (int a = 0; a < n; a++) g += J[a * n + c] * r[a]; g2 += g * g; }
                    o[0] = n; o[1] = tr; o[2] = g2;
                }
            }
        }
    free(x0); free(J); free(r); free(st);
    return NULL;
}
long vo_optimization_many(int n, const viwb_problem *problems, const double *const *states, const int32_t *flags,
                          const viwb_options *opt, int threads, int repeat, double *const *out_states, int32_t *out_iters, double *out_prior);
long vo_optimization_throughput(int n, const viwb_problem *problems, const double *const *states, const int32_t *flags,
                                const viwb_options *opt, int threads, int repeat) {
    return vo_optimization_many(n, problems, states, flags, opt, threads, repeat, NULL, NULL, NULL);
}
long vo_optimization_many(int n, const viwb_problem *problems, const double *const *states, const int32_t *flags,
                          const viwb_options *opt, int threads, int repeat, double *const *out_states, int32_t *out_iters, double *out_prior) {
    if (threads < 1) threads = 1;
    /* the per-solve work buffers (MBs) would otherwise be mmap'ed/unmapped on every call and the threads would serialise on
     * the process' mm lock: keep them in the per-thread malloc arenas (gives the CPU arm its best case) */
    mallopt(M_MMAP_THRESHOLD, 1 << 30); mallopt(M_TRIM_THRESHOLD, 1 << 30); mallopt(M_TOP_PAD, 64 << 20);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    vo_job *jobs = (vo_job *)calloc(threads, sizeof(vo_job));
    for (int t = 0; t < threads; t++) {
        jobs[t].pb = problems; jobs[t].states = states; jobs[t].flags = flags; jobs[t].opt = opt; jobs[t].n = n; jobs[t].repeat = repeat; jobs[t].tid = t; jobs[t].nthreads = threads;
        jobs[t].out_states = out_states; jobs[t].out_iters = out_iters; jobs[t].out_prior = out_prior;
        pthread_create(&th[t], NULL, vo_worker, &jobs[t]);
    }
    long done = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); done += jobs[t].done; }
    free(th); free(jobs);
    return done;
}

/*
 * vo_solver.c -- CPU oracle, part 2: the ceres::Solve(DENSE_SCHUR + DOGLEG) restatement.
 * TEST INFRASTRUCTURE ONLY (see viw_oracle.h).
 *
 * Third-party arithmetic restated here (NOT in /root/reference; ceres-solver 1.14.0 per README.md:41):
 *   Program::RemoveFixedBlocks, ResidualBlock::Evaluate + Corrector (corrector.cc),
 *   TrustRegionMinimizer::Minimize (trust_region_minimizer.cc), DoglegStrategy (dogleg_strategy.cc,
 *   TRADITIONAL_DOGLEG), DenseSchurComplementSolver (schur_complement_solver.cc, Eigen LLT).
 * Call site in the reference: estimator/estimator.cpp:1388-1658.
 */
#include "viw_oracle.h"
#include "vo_math.h"
#include <stdlib.h>
#include <float.h>

/* ------------------------------------------------------------------ program */
#define MAX_SLOTS 32   /* parameter slots of one residual block (prior: up to 32 kept blocks) */
typedef struct {
    int type;            /* viwb_factor_type, or -1 for the prior */
    int nrows, nslots;
    int block[MAX_SLOTS];    /* block id (fixed id, or 32+k) */
    int col[MAX_SLOTS];      /* column offset in the reduced tangent vector, -1 if constant */
    int tsize[MAX_SLOTS];
    const double *consts;
    int row0;
    size_t jac0[MAX_SLOTS];  /* offset of the local Jacobian (nrows x tsize, row-major) in prog.jac */
    int has_loss;
} rblock_t;

typedef struct {
    const viwb_problem *pb;
    int fixed_col[VIWB_NUM_FIXED_BLOCKS];   /* -1 if inactive */
    int n_active_fixed, active_fixed[VIWB_NUM_FIXED_BLOCKS];
    int nf;               /* tangent columns of the active fixed blocks */
    int nlm;              /* landmarks (all active) */
    int *lm_col;          /* column of landmark k (nf + k') or -1 */
    int ncols, nrows, namb;
    int nrb; rblock_t *rb;
    double *jac; size_t jac_size;
    double *residuals;
    double fixed_cost;
} program_t;

static int block_gsize(int b) { return b < VIWB_NUM_FIXED_BLOCKS ? viwb_block_size(b) : 1; }
static int block_tsize(int b) { return b < VIWB_NUM_FIXED_BLOCKS ? viwb_block_tsize(b) : 1; }
static int block_soff(int b) { return b < VIWB_NUM_FIXED_BLOCKS ? viwb_block_offset(b) : VIWB_STATE_FIXED + (b - VIWB_NUM_FIXED_BLOCKS); }

static void rb_set(rblock_t *r, int type, int nrows, int nslots, const int *blocks, const double *consts, int has_loss) {
    r->type = type; r->nrows = nrows; r->nslots = nslots; r->consts = consts; r->has_loss = has_loss;
    for (int i = 0; i < nslots; i++) r->block[i] = blocks[i];
}

static int program_build(program_t *P, const viwb_problem *pb) {
    memset(P, 0, sizeof *P);
    P->pb = pb;
    int has_prior = pb->prior && pb->prior->valid;
    P->nrb = (has_prior ? 1 : 0) + pb->num_imu + pb->num_wheel + pb->num_plane + pb->num_vis;
    P->rb = (rblock_t *)calloc(P->nrb > 0 ? P->nrb : 1, sizeof(rblock_t));
    int k = 0, blocks[MAX_SLOTS];
    /* order of AddResidualBlock calls: prior, IMU, wheel, plane, visual (estimator.cpp:1521-1638) */
    if (has_prior) {
        if (pb->prior->num_blocks > MAX_SLOTS) return VIWB_ERR_INVALID;
        for (int i = 0; i < pb->prior->num_blocks; i++) blocks[i] = pb->prior->block_id[i];
        rb_set(&P->rb[k++], -1, pb->prior->n, pb->prior->num_blocks, blocks, NULL, 0);
    }
    for (int i = 0; i < pb->num_imu; i++) {
        int fi = pb->imu_frame_i[i], fj = pb->imu_frame_j[i];
        int b[4] = {VIWB_BLK_POSE0 + fi, VIWB_BLK_SPEEDBIAS0 + fi, VIWB_BLK_POSE0 + fj, VIWB_BLK_SPEEDBIAS0 + fj};
        rb_set(&P->rb[k++], VIWB_F_IMU, 15, 4, b, pb->imu_data + (size_t)i * VIWB_IMU_DOUBLES, 0);
    }
    for (int i = 0; i < pb->num_wheel; i++) {
        int fi = pb->wheel_frame_i[i], fj = pb->wheel_frame_j[i];
        int b[7] = {VIWB_BLK_POSE0 + fi, VIWB_BLK_POSE0 + fj, VIWB_BLK_EX_WHEEL, VIWB_BLK_SX, VIWB_BLK_SY, VIWB_BLK_SW, VIWB_BLK_TD_WHEEL};
        rb_set(&P->rb[k++], VIWB_F_WHEEL, 6, 7, b, pb->wheel_data + (size_t)i * VIWB_WHEEL_DOUBLES, 0);
    }
    for (int i = 0; i < pb->num_plane; i++) {
        int b[4] = {VIWB_BLK_POSE0 + pb->plane_frame[i], VIWB_BLK_EX_WHEEL, VIWB_BLK_PLANE_R, VIWB_BLK_PLANE_Z};
        rb_set(&P->rb[k++], VIWB_F_PLANE, 3, 4, b, NULL, 0);
    }
    for (int i = 0; i < pb->num_vis; i++) {
        int fi = pb->vis_frame_i[i], fj = pb->vis_frame_j[i], lm = VIWB_BLK_LANDMARK0 + pb->vis_landmark[i];
        const double *c = pb->vis_obs + (size_t)i * VIWB_VIS_OBS_DOUBLES;
        if (pb->vis_type[i] == VIWB_F_PROJ_2F1C) {
            int b[5] = {VIWB_BLK_POSE0 + fi, VIWB_BLK_POSE0 + fj, VIWB_BLK_EX_POSE0, lm, VIWB_BLK_TD};
            rb_set(&P->rb[k++], VIWB_F_PROJ_2F1C, 2, 5, b, c, 1);
        } else if (pb->vis_type[i] == VIWB_F_PROJ_2F2C) {
            int b[6] = {VIWB_BLK_POSE0 + fi, VIWB_BLK_POSE0 + fj, VIWB_BLK_EX_POSE0, VIWB_BLK_EX_POSE1, lm, VIWB_BLK_TD};
            rb_set(&P->rb[k++], VIWB_F_PROJ_2F2C, 2, 6, b, c, 1);
        } else if (pb->vis_type[i] == VIWB_F_PROJ_1F2C) {
            int b[4] = {VIWB_BLK_EX_POSE0, VIWB_BLK_EX_POSE1, lm, VIWB_BLK_TD};
            rb_set(&P->rb[k++], VIWB_F_PROJ_1F2C, 2, 4, b, c, 1);
        } else return VIWB_ERR_INVALID;
    }
    /* Program::RemoveFixedBlocks: a block is a variable iff it is not constant and some residual depends on it */
    int referenced[VIWB_NUM_FIXED_BLOCKS] = {0};
    P->nlm = pb->num_landmarks;
    int *lm_ref = (int *)calloc(P->nlm > 0 ? P->nlm : 1, sizeof(int));
    for (int r = 0; r < P->nrb; r++) for (int s = 0; s < P->rb[r].nslots; s++) {
        int b = P->rb[r].block[s];
        if (b < VIWB_NUM_FIXED_BLOCKS) referenced[b] = 1;
        else { if (b - VIWB_NUM_FIXED_BLOCKS >= P->nlm) { free(lm_ref); return VIWB_ERR_INVALID; } lm_ref[b - VIWB_NUM_FIXED_BLOCKS] = 1; }
    }
    int col = 0, amb = 0;
    for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) {
        int active = (pb->block_flags[b] & VIWB_BLOCK_PRESENT) && !(pb->block_flags[b] & VIWB_BLOCK_CONSTANT) && referenced[b];
        if (active) { P->fixed_col[b] = col; col += viwb_block_tsize(b); amb += viwb_block_size(b); P->active_fixed[P->n_active_fixed++] = b; }
        else P->fixed_col[b] = -1;
    }
    P->nf = col;
    P->lm_col = (int *)malloc(sizeof(int) * (P->nlm > 0 ? P->nlm : 1));
    for (int i = 0; i < P->nlm; i++) { if (lm_ref[i]) { P->lm_col[i] = col++; amb++; } else P->lm_col[i] = -1; }
    free(lm_ref);
    P->ncols = col; P->namb = amb;
    int row = 0; size_t joff = 0;
    for (int r = 0; r < P->nrb; r++) {
        rblock_t *R = &P->rb[r];
        R->row0 = row; row += R->nrows;
        for (int s = 0; s < R->nslots; s++) {
            int b = R->block[s];
            R->tsize[s] = block_tsize(b);
            R->col[s] = b < VIWB_NUM_FIXED_BLOCKS ? P->fixed_col[b] : P->lm_col[b - VIWB_NUM_FIXED_BLOCKS];
            R->jac0[s] = joff; joff += (size_t)R->nrows * R->tsize[s];
        }
    }
    P->nrows = row; P->jac_size = joff;
    P->jac = (double *)malloc(sizeof(double) * (joff > 0 ? joff : 1));
    P->residuals = (double *)malloc(sizeof(double) * (row > 0 ? row : 1));
    return 0;
}
static void program_free(program_t *P) { free(P->rb); free(P->lm_col); free(P->jac); free(P->residuals); }

/* ------------------------------------------------------------------ evaluator */
/* ResidualBlock::Evaluate + Corrector (ceres corrector.cc; in-tree twin marginalization_factor.cpp:46-77).
 * Fills prog.residuals (corrected) and, if want_jac, prog.jac (local, corrected, unscaled). */
static int program_evaluate(program_t *P, const double *state, int want_jac, double *cost_out) {
    const viwb_problem *pb = P->pb;
    double cost = 0, fixed_cost = 0;
    double gj[MAX_SLOTS > 7 ? 7 : MAX_SLOTS][15 * 9];   /* global Jacobians of one analytic factor (<= 7 slots, <= 15x9) */
    double *prior_jac = NULL;
    for (int r = 0; r < P->nrb; r++) {
        rblock_t *R = &P->rb[r];
        double *res = P->residuals + R->row0;
        int any_active = 0;
        for (int s = 0; s < R->nslots; s++) if (R->col[s] >= 0) any_active = 1;
        int need_jac = want_jac && any_active;
        if (R->type < 0) {
            if (need_jac && !prior_jac) prior_jac = (double *)malloc(sizeof(double) * pb->prior->n * VIWB_STATE_FIXED);
            vo_prior_evaluate(pb->prior, state, res, need_jac ? prior_jac : NULL);
            if (need_jac) for (int s = 0; s < R->nslots; s++) {
                if (R->col[s] < 0) continue;
                int off = viwb_block_offset(R->block[s]), ts = R->tsize[s];
                double *J = P->jac + R->jac0[s];
                for (int i = 0; i < R->nrows; i++) for (int c = 0; c < ts; c++) J[i * ts + c] = prior_jac[i * VIWB_STATE_FIXED + off + c];
            }
        } else {
            const double *params[7]; double *jp[7];
            for (int s = 0; s < R->nslots; s++) { params[s] = state + block_soff(R->block[s]); jp[s] = (need_jac && R->col[s] >= 0) ? gj[s] : NULL; }
            int rc = vo_factor_evaluate(R->type, &pb->globals, R->consts, params, res, need_jac ? jp : NULL);
            if (rc) { free(prior_jac); return rc; }
            if (need_jac) for (int s = 0; s < R->nslots; s++) {
                if (R->col[s] < 0) continue;
                int gs = block_gsize(R->block[s]), ts = R->tsize[s];
                double *J = P->jac + R->jac0[s];
                /* local = global * ComputeJacobian, ComputeJacobian = [I;0] (pose_local_parameterization.cpp:28-36) */
                for (int i = 0; i < R->nrows; i++) for (int c = 0; c < ts; c++) J[i * ts + c] = gj[s][i * gs + c];
            }
        }
        double sq = 0; for (int i = 0; i < R->nrows; i++) sq += res[i] * res[i];
        double c;
        if (!R->has_loss) c = 0.5 * sq;
        else {
            double rho[3]; vo_huber(pb->globals.huber_delta, sq, rho);
            c = 0.5 * rho[0];
            double sqrt_rho1 = sqrt(rho[1]), residual_scaling, alpha_sq_norm;
            if (sq == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
            else {
                double D = 1.0 + 2.0 * sq * rho[2] / rho[1], alpha = 1.0 - sqrt(D);
                residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / sq;
            }
            if (need_jac) for (int s = 0; s < R->nslots; s++) {
                if (R->col[s] < 0) continue;
                int ts = R->tsize[s]; double *J = P->jac + R->jac0[s];
                if (alpha_sq_norm == 0.0) { for (int i = 0; i < R->nrows * ts; i++) J[i] *= sqrt_rho1; }
                else for (int cidx = 0; cidx < ts; cidx++) {
                    double rtj = 0; for (int i = 0; i < R->nrows; i++) rtj += res[i] * J[i * ts + cidx];
                    for (int i = 0; i < R->nrows; i++) J[i * ts + cidx] = sqrt_rho1 * (J[i * ts + cidx] - alpha_sq_norm * res[i] * rtj);
                }
            }
            for (int i = 0; i < R->nrows; i++) res[i] *= residual_scaling;
        }
        if (any_active) cost += c; else fixed_cost += c;
    }
    free(prior_jac);
    P->fixed_cost = fixed_cost;
    *cost_out = cost;
    return 0;
}

/* y += J x  /  y += J^T x over active columns (skips constant-only blocks' rows too) */
static void jac_right_multiply(const program_t *P, const double *x, double *y) {
    for (int r = 0; r < P->nrb; r++) { const rblock_t *R = &P->rb[r];
        for (int s = 0; s < R->nslots; s++) { if (R->col[s] < 0) continue;
            int ts = R->tsize[s]; const double *J = P->jac + R->jac0[s];
            for (int i = 0; i < R->nrows; i++) { double a = 0; for (int c = 0; c < ts; c++) a += J[i * ts + c] * x[R->col[s] + c]; y[R->row0 + i] += a; } } }
}
static void jac_left_multiply(const program_t *P, const double *x, double *y) {
    for (int r = 0; r < P->nrb; r++) { const rblock_t *R = &P->rb[r];
        for (int s = 0; s < R->nslots; s++) { if (R->col[s] < 0) continue;
            int ts = R->tsize[s]; const double *J = P->jac + R->jac0[s];
            for (int i = 0; i < R->nrows; i++) { double xi = x[R->row0 + i]; for (int c = 0; c < ts; c++) y[R->col[s] + c] += J[i * ts + c] * xi; } } }
}
static void jac_sq_col_norm(const program_t *P, double *d) {
    for (int c = 0; c < P->ncols; c++) d[c] = 0;
    for (int r = 0; r < P->nrb; r++) { const rblock_t *R = &P->rb[r];
        for (int s = 0; s < R->nslots; s++) { if (R->col[s] < 0) continue;
            int ts = R->tsize[s]; const double *J = P->jac + R->jac0[s];
            for (int i = 0; i < R->nrows; i++) for (int c = 0; c < ts; c++) d[R->col[s] + c] += J[i * ts + c] * J[i * ts + c]; } }
}
static void jac_scale_cols(program_t *P, const double *sc) {
    for (int r = 0; r < P->nrb; r++) { rblock_t *R = &P->rb[r];
        for (int s = 0; s < R->nslots; s++) { if (R->col[s] < 0) continue;
            int ts = R->tsize[s]; double *J = P->jac + R->jac0[s];
            for (int i = 0; i < R->nrows; i++) for (int c = 0; c < ts; c++) J[i * ts + c] *= sc[R->col[s] + c]; } }
}

/* ------------------------------------------------------------------ manifold */
/* PoseLocalParameterization::Plus (pose_local_parameterization.cpp:12-27), PoseSubsetParameterization::Plus
 * (pose_subset_parameterization.cpp:27-53), OrientationSubsetParameterization::Plus (orientation_subset_
 * parameterization.cpp:27-44); Euclidean blocks: x + delta. */
static void block_plus(int b, unsigned mask, const double *x, const double *delta, double *out) {
    int gs = block_gsize(b);
    if (gs == 7) {
        double d[6]; for (int i = 0; i < 6; i++) d[i] = (mask >> i) & 1u ? 0.0 : delta[i];
        for (int i = 0; i < 3; i++) out[i] = x[i] + d[i];
        double dq[4], q[4]; q_delta(dq, d + 3); q_mul(q, x + 3, dq); q_normalize(q);
        memcpy(out + 3, q, sizeof q);
    } else if (gs == 4) {
        double d[3]; for (int i = 0; i < 3; i++) d[i] = (mask >> i) & 1u ? 0.0 : delta[i];
        double dq[4], q[4]; q_delta(dq, d); q_mul(q, x, dq); q_normalize(q);
        memcpy(out, q, sizeof q);
    } else for (int i = 0; i < gs; i++) out[i] = x[i] + delta[i];
}
/* reduced-vector Plus: out = state with the active blocks replaced by x (+) delta */
static void program_plus(const program_t *P, const double *state, const double *delta, double *out) {
    int n = VIWB_STATE_FIXED + P->nlm;
    if (out != state) memcpy(out, state, sizeof(double) * n);
    for (int i = 0; i < P->n_active_fixed; i++) {
        int b = P->active_fixed[i]; double tmp[9];
        block_plus(b, P->pb->subset_mask[b], state + viwb_block_offset(b), delta + P->fixed_col[b], tmp);
        memcpy(out + viwb_block_offset(b), tmp, sizeof(double) * viwb_block_size(b));
    }
    for (int k = 0; k < P->nlm; k++) if (P->lm_col[k] >= 0) out[VIWB_STATE_FIXED + k] = state[VIWB_STATE_FIXED + k] + delta[P->lm_col[k]];
}
static double program_amb_norm(const program_t *P, const double *a, const double *b /* may be NULL */, int inf_norm) {
    double acc = 0;
    for (int i = 0; i < P->n_active_fixed; i++) {
        int blk = P->active_fixed[i], off = viwb_block_offset(blk);
        for (int k = 0; k < viwb_block_size(blk); k++) { double d = a[off + k] - (b ? b[off + k] : 0); if (inf_norm) { if (fabs(d) > acc) acc = fabs(d); } else acc += d * d; }
    }
    for (int k = 0; k < P->nlm; k++) if (P->lm_col[k] >= 0) { double d = a[VIWB_STATE_FIXED + k] - (b ? b[VIWB_STATE_FIXED + k] : 0); if (inf_norm) { if (fabs(d) > acc) acc = fabs(d); } else acc += d * d; }
    return inf_norm ? acc : sqrt(acc);
}
void vo_state_plus(const viwb_problem *pb, const double *state, const double *delta, double *out) {
    int n = VIWB_STATE_FIXED + pb->num_landmarks;
    double *tmp = (double *)malloc(sizeof(double) * n);
    memcpy(tmp, state, sizeof(double) * n);
    for (int b = 0; b < VIWB_NUM_FIXED_BLOCKS; b++) {
        double t[9]; block_plus(b, pb->subset_mask[b], state + viwb_block_offset(b), delta + viwb_block_toffset(b), t);
        memcpy(tmp + viwb_block_offset(b), t, sizeof(double) * viwb_block_size(b));
    }
    for (int k = 0; k < pb->num_landmarks; k++) tmp[VIWB_STATE_FIXED + k] = state[VIWB_STATE_FIXED + k] + delta[VIWB_TANGENT_FIXED + k];
    memcpy(out, tmp, sizeof(double) * n);
    free(tmp);
}

/* ------------------------------------------------------------------ dense Schur linear solver */
typedef struct {
    int nf, ne;
    double *H, *W, *h, *gf, *ge, *S, *rhs, *yf;    /* H nf*nf, W ne*nf (row k = column of landmark k), h ne */
    int *lm_nblk, *lm_blk;                         /* touched fixed column ranges per landmark: (col,size) pairs, <= 16 */
} schur_t;
static void schur_alloc(schur_t *s, int nf, int ne) {
    s->nf = nf; s->ne = ne;
    size_t nf1 = nf > 0 ? nf : 1, ne1 = ne > 0 ? ne : 1;
    s->H = (double *)malloc(sizeof(double) * nf1 * nf1); s->W = (double *)malloc(sizeof(double) * ne1 * nf1);
    s->h = (double *)malloc(sizeof(double) * ne1); s->gf = (double *)malloc(sizeof(double) * nf1); s->ge = (double *)malloc(sizeof(double) * ne1);
    s->S = (double *)malloc(sizeof(double) * nf1 * nf1); s->rhs = (double *)malloc(sizeof(double) * nf1); s->yf = (double *)malloc(sizeof(double) * nf1);
    s->lm_nblk = (int *)calloc(ne1, sizeof(int)); s->lm_blk = (int *)malloc(sizeof(int) * ne1 * 32);
}
static void schur_free(schur_t *s) { free(s->H); free(s->W); free(s->h); free(s->gf); free(s->ge); free(s->S); free(s->rhs); free(s->yf); free(s->lm_nblk); free(s->lm_blk); }

/* normal equations of the (scaled) block-sparse Jacobian: H = Jf^T Jf, W = Jf^T Je, h = diag(Je^T Je), g = J^T r */
static void schur_accumulate(schur_t *s, const program_t *P) {
    int nf = s->nf, ne = s->ne;
    memset(s->H, 0, sizeof(double) * nf * nf); memset(s->W, 0, sizeof(double) * (size_t)ne * nf);
    memset(s->h, 0, sizeof(double) * ne); memset(s->gf, 0, sizeof(double) * nf); memset(s->ge, 0, sizeof(double) * ne);
    memset(s->lm_nblk, 0, sizeof(int) * ne);
    for (int r = 0; r < P->nrb; r++) {
        const rblock_t *R = &P->rb[r];
        const double *res = P->residuals + R->row0;
        for (int a = 0; a < R->nslots; a++) {
            if (R->col[a] < 0) continue;
            int ca = R->col[a], ta = R->tsize[a]; const double *Ja = P->jac + R->jac0[a];
            /* gradient */
            for (int i = 0; i < R->nrows; i++) for (int c = 0; c < ta; c++) {
                if (ca < nf) s->gf[ca + c] += Ja[i * ta + c] * res[i]; else s->ge[ca - nf] += Ja[i * ta + c] * res[i];
            }
            for (int b = a; b < R->nslots; b++) {
                if (R->col[b] < 0) continue;
                int cb = R->col[b], tb = R->tsize[b]; const double *Jb = P->jac + R->jac0[b];
                if (ca < nf && cb < nf) {
                    for (int c = 0; c < ta; c++) for (int d = 0; d < tb; d++) {
                        double v = 0; for (int i = 0; i < R->nrows; i++) v += Ja[i * ta + c] * Jb[i * tb + d];
                        s->H[(ca + c) * nf + cb + d] += v;
                        if (a != b) s->H[(cb + d) * nf + ca + c] += v;
                    }
                } else if (ca >= nf && cb >= nf) { /* same landmark (a == b) */
                    double v = 0; for (int i = 0; i < R->nrows; i++) v += Ja[i] * Jb[i];
                    s->h[ca - nf] += v;
                } else {
                    int cf = ca < nf ? ca : cb, tf = ca < nf ? ta : tb, k = (ca < nf ? cb : ca) - nf;
                    const double *Jf = ca < nf ? Ja : Jb, *Je = ca < nf ? Jb : Ja;
                    double *Wk = s->W + (size_t)k * nf;
                    for (int c = 0; c < tf; c++) { double v = 0; for (int i = 0; i < R->nrows; i++) v += Jf[i * tf + c] * Je[i]; Wk[cf + c] += v; }
                    int found = 0; for (int q = 0; q < s->lm_nblk[k]; q++) if (s->lm_blk[k * 32 + 2 * q] == cf) found = 1;
                    if (!found && s->lm_nblk[k] < 16) { s->lm_blk[k * 32 + 2 * s->lm_nblk[k]] = cf; s->lm_blk[k * 32 + 2 * s->lm_nblk[k] + 1] = tf; s->lm_nblk[k]++; }
                }
            }
        }
    }
}
/* solve (J^T J + D^2) y = J^T r by eliminating the landmark block; returns 0 ok, 1 = LINEAR_SOLVER_FAILURE */
static int schur_solve(schur_t *s, const double *D, double *y) {
    int nf = s->nf, ne = s->ne;
    memcpy(s->S, s->H, sizeof(double) * nf * nf); memcpy(s->rhs, s->gf, sizeof(double) * nf);
    for (int i = 0; i < nf; i++) s->S[i * nf + i] += D[i] * D[i];
    for (int k = 0; k < ne; k++) {
        double hk = s->h[k] + D[nf + k] * D[nf + k], inv = 1.0 / hk;
        const double *Wk = s->W + (size_t)k * nf;
        int nb = s->lm_nblk[k]; const int *bl = s->lm_blk + k * 32;
        for (int a = 0; a < nb; a++) for (int c = 0; c < bl[2 * a + 1]; c++) {
            int row = bl[2 * a] + c; double wa = Wk[row] * inv;
            s->rhs[row] -= wa * s->ge[k];
            for (int b = 0; b < nb; b++) for (int d = 0; d < bl[2 * b + 1]; d++) s->S[row * nf + bl[2 * b] + d] -= wa * Wk[bl[2 * b] + d];
        }
    }
    /* Eigen LLT (selfadjointView<Upper>().llt()): fails on a non-positive pivot */
    double *L = s->S;
    for (int j = 0; j < nf; j++) {
        double d = L[j * nf + j];
        for (int k = 0; k < j; k++) d -= L[j * nf + k] * L[j * nf + k];
        if (!(d > 0)) return 1;
        d = sqrt(d); L[j * nf + j] = d;
        for (int i = j + 1; i < nf; i++) {
            double v = L[i * nf + j];
            for (int k = 0; k < j; k++) v -= L[i * nf + k] * L[j * nf + k];
            L[i * nf + j] = v / d;
        }
    }
    for (int i = 0; i < nf; i++) { double v = s->rhs[i]; for (int k = 0; k < i; k++) v -= L[i * nf + k] * s->yf[k]; s->yf[i] = v / L[i * nf + i]; }
    for (int i = nf - 1; i >= 0; i--) { double v = s->yf[i]; for (int k = i + 1; k < nf; k++) v -= L[k * nf + i] * s->yf[k]; s->yf[i] = v / L[i * nf + i]; }
    for (int i = 0; i < nf; i++) y[i] = s->yf[i];
    for (int k = 0; k < ne; k++) {
        double hk = s->h[k] + D[nf + k] * D[nf + k], v = s->ge[k];
        const double *Wk = s->W + (size_t)k * nf;
        int nb = s->lm_nblk[k]; const int *bl = s->lm_blk + k * 32;
        for (int a = 0; a < nb; a++) for (int c = 0; c < bl[2 * a + 1]; c++) v -= Wk[bl[2 * a] + c] * s->yf[bl[2 * a] + c];
        y[nf + k] = v / hk;
    }
    for (int i = 0; i < nf + ne; i++) if (!isfinite(y[i])) return 1;
    return 0;
}

/* ------------------------------------------------------------------ trust region minimizer */
static void trace_push(vo_trace *t, const vo_trace_entry *e) { if (t && t->count < VO_MAX_TRACE) t->e[t->count++] = *e; }

int vo_window_solve(const viwb_problem *pb, double *state, const viwb_options *opt, viwb_summary *sum, vo_trace *trace) {
    program_t P; int rc = program_build(&P, pb);
    if (rc) { program_free(&P); return rc; }
    if (trace) trace->count = 0;
    const int n = P.ncols, nstate = VIWB_STATE_FIXED + P.nlm, m = P.nrows;
    double *x = (double *)malloc(sizeof(double) * nstate * 2), *cand = x + nstate;
    memcpy(x, state, sizeof(double) * nstate);
    double *scale = (double *)malloc(sizeof(double) * (n + 1) * 9);
    double *gradient = scale + (n + 1), *diag = gradient + (n + 1), *sgrad = diag + (n + 1), *gn = sgrad + (n + 1);
    double *step = gn + (n + 1), *delta = step + (n + 1), *tmpn = delta + (n + 1), *lmdiag = tmpn + (n + 1);
    double *tmpm = (double *)malloc(sizeof(double) * (m + 1));
    schur_t S; schur_alloc(&S, P.nf, n - P.nf);
    memset(sum, 0, sizeof *sum);

    /* DoglegStrategy state (dogleg_strategy.cc) */
    double radius = opt->initial_trust_region_radius, mu = 1e-8, alpha = 0, dogleg_step_norm = 0;
    const double min_mu = 1e-8, max_mu = 1.0, mu_increase = 10.0;
    int reuse = 0, num_invalid = 0, num_linear = 0;
    double x_cost = 0, x_norm = program_amb_norm(&P, x, NULL, 0);
    int iteration = 0, successful_steps = 0, term = VIWB_NO_CONVERGENCE;
    vo_trace_entry te; memset(&te, 0, sizeof te);
    double gradient_max_norm = 0;

#define EVAL_GRADIENT_AND_JACOBIAN(first) do { \
        rc = program_evaluate(&P, x, 1, &x_cost); if (rc) goto done; \
        for (int i_ = 0; i_ < n; i_++) gradient[i_] = 0; \
        jac_left_multiply(&P, P.residuals, gradient);                      /* g = J^T r, unscaled */ \
        for (int i_ = 0; i_ < n; i_++) tmpn[i_] = -gradient[i_]; \
        program_plus(&P, x, tmpn, cand); \
        gradient_max_norm = program_amb_norm(&P, x, cand, 1); \
        if (opt->jacobi_scaling) { \
            if (first) { jac_sq_col_norm(&P, scale); for (int i_ = 0; i_ < n; i_++) scale[i_] = 1.0 / (1.0 + sqrt(scale[i_])); } \
            jac_scale_cols(&P, scale); \
        } else if (first) for (int i_ = 0; i_ < n; i_++) scale[i_] = 1.0; \
    } while (0)

    /* IterationZero */
    EVAL_GRADIENT_AND_JACOBIAN(1);
    sum->initial_cost = x_cost + P.fixed_cost;
    te.iteration = 0; te.step_is_valid = 1; te.step_is_successful = 1; te.cost = x_cost; te.radius = radius; te.mu = mu;
    te.gradient_max_norm = gradient_max_norm; te.x_norm = x_norm;
    trace_push(trace, &te); sum->num_iterations = 1;
    if (gradient_max_norm <= opt->gradient_tolerance) { term = VIWB_CONVERGENCE; goto done; }
    if (radius <= opt->min_trust_region_radius) { term = VIWB_CONVERGENCE; goto done; }

    for (;;) {
        if (iteration >= opt->max_num_iterations) { term = VIWB_NO_CONVERGENCE; break; }
        iteration++;
        memset(&te, 0, sizeof te); te.iteration = iteration;
        /* ---- DoglegStrategy::ComputeStep ---- */
        int ls_failure = 0;
        te.reused = reuse;
        if (!reuse) {
            reuse = 1;
            jac_sq_col_norm(&P, diag);
            for (int i = 0; i < n; i++) { double d = diag[i]; if (d < opt->min_lm_diagonal) d = opt->min_lm_diagonal; if (d > opt->max_lm_diagonal) d = opt->max_lm_diagonal; diag[i] = sqrt(d); }
            /* ComputeGradient: g = J^T r ./ D */
            for (int i = 0; i < n; i++) sgrad[i] = 0;
            jac_left_multiply(&P, P.residuals, sgrad);
            for (int i = 0; i < n; i++) sgrad[i] /= diag[i];
            /* ComputeCauchyPoint */
            for (int i = 0; i < n; i++) tmpn[i] = sgrad[i] / diag[i];
            for (int i = 0; i < m; i++) tmpm[i] = 0;
            jac_right_multiply(&P, tmpn, tmpm);
            double g2 = 0, Jg2 = 0; for (int i = 0; i < n; i++) g2 += sgrad[i] * sgrad[i]; for (int i = 0; i < m; i++) Jg2 += tmpm[i] * tmpm[i];
            alpha = g2 / Jg2;
            /* ComputeGaussNewtonStep */
            schur_accumulate(&S, &P);
            ls_failure = 1;
            while (mu < max_mu) {
                for (int i = 0; i < n; i++) lmdiag[i] = diag[i] * sqrt(mu);
                num_linear++;
                if (schur_solve(&S, lmdiag, gn)) { mu *= mu_increase; ls_failure = 1; continue; }
                ls_failure = 0; break;
            }
            if (!ls_failure) for (int i = 0; i < n; i++) gn[i] *= -diag[i];
        }
        int step_valid = 0; double model_cost_change = 0;
        if (!ls_failure) {
            /* ComputeTraditionalDoglegStep */
            double gnorm = 0, gnnorm = 0; for (int i = 0; i < n; i++) { gnorm += sgrad[i] * sgrad[i]; gnnorm += gn[i] * gn[i]; }
            gnorm = sqrt(gnorm); gnnorm = sqrt(gnnorm);
            if (gnnorm <= radius) { for (int i = 0; i < n; i++) step[i] = gn[i]; dogleg_step_norm = gnnorm; }
            else if (gnorm * alpha >= radius) { for (int i = 0; i < n; i++) step[i] = -(radius / gnorm) * sgrad[i]; dogleg_step_norm = radius; }
            else {
                double gdot = 0; for (int i = 0; i < n; i++) gdot += sgrad[i] * gn[i];
                double b_dot_a = -alpha * gdot, a_sq = pow(alpha * gnorm, 2.0);
                double bma_sq = a_sq - 2 * b_dot_a + pow(gnnorm, 2);
                double c = b_dot_a - a_sq, d = sqrt(c * c + bma_sq * (pow(radius, 2.0) - a_sq));
                double beta = (c <= 0) ? (d - c) / bma_sq : (radius * radius - a_sq) / (d + c);
                double nn = 0;
                for (int i = 0; i < n; i++) { step[i] = (-alpha * (1.0 - beta)) * sgrad[i] + beta * gn[i]; nn += step[i] * step[i]; }
                dogleg_step_norm = sqrt(nn);
            }
            for (int i = 0; i < n; i++) step[i] /= diag[i];
            /* model_cost_change = -(J step)^T (r + J step / 2) */
            for (int i = 0; i < m; i++) tmpm[i] = 0;
            jac_right_multiply(&P, step, tmpm);
            for (int i = 0; i < m; i++) model_cost_change -= tmpm[i] * (P.residuals[i] + tmpm[i] / 2.0);
            step_valid = model_cost_change > 0.0;
        }
        te.step_is_valid = step_valid; te.model_cost_change = model_cost_change; te.alpha = alpha; te.dogleg_step_norm = dogleg_step_norm;
        if (!step_valid) {
            /* HandleInvalidStep */
            if (++num_invalid >= opt->max_num_consecutive_invalid_steps) { term = VIWB_FAILURE; break; }
            mu *= mu_increase; reuse = 0;                                   /* DoglegStrategy::StepIsInvalid */
            te.cost = x_cost; te.radius = radius; te.mu = mu; te.gradient_max_norm = gradient_max_norm; te.x_norm = x_norm;
            trace_push(trace, &te); sum->num_iterations++;
            if (radius <= opt->min_trust_region_radius) { term = VIWB_CONVERGENCE; break; }
            continue;
        }
        num_invalid = 0;
        for (int i = 0; i < n; i++) delta[i] = step[i] * scale[i];
        /* ComputeCandidatePointAndEvaluateCost */
        program_plus(&P, x, delta, cand);
        double cand_cost; { double c_; program_t *PP = &P; rc = program_evaluate(PP, cand, 0, &c_); if (rc) goto done; cand_cost = c_; }
        if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
        /* the evaluation above overwrote P.residuals with the candidate's; restore lazily below if rejected */
        double step_norm = program_amb_norm(&P, x, cand, 0);
        te.step_norm = step_norm; te.x_norm = x_norm;
        if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) { term = VIWB_CONVERGENCE; te.cost = x_cost; trace_push(trace, &te); break; }
        double cost_change = x_cost - cand_cost;
        te.cost_change = cost_change;
        if (fabs(cost_change) <= opt->function_tolerance * x_cost) { term = VIWB_CONVERGENCE; te.cost = x_cost; trace_push(trace, &te); break; }
        double relative_decrease = cost_change / model_cost_change;
        te.relative_decrease = relative_decrease;
        if (relative_decrease > opt->min_relative_decrease) {
            /* HandleSuccessfulStep */
            memcpy(x, cand, sizeof(double) * nstate); x_norm = program_amb_norm(&P, x, NULL, 0);
            EVAL_GRADIENT_AND_JACOBIAN(0);
            te.step_is_successful = 1; successful_steps++;
            /* DoglegStrategy::StepAccepted */
            if (relative_decrease < 0.25) radius *= 0.5;
            if (relative_decrease > 0.75) { double r3 = 3.0 * dogleg_step_norm; if (r3 > radius) radius = r3; }
            /* (DoglegStrategy::StepAccepted does not clamp to max_trust_region_radius; only LevenbergMarquardtStrategy does) */
            mu = 2.0 * mu / mu_increase; if (mu < min_mu) mu = min_mu;
            reuse = 0;
        } else {
            /* HandleUnsuccessfulStep; DoglegStrategy::StepRejected */
            radius *= 0.5; reuse = 1;
            /* restore residuals at x (Ceres keeps residuals_ of x; cost-only evaluation does not touch them) */
            { double c_; rc = program_evaluate(&P, x, 0, &c_); if (rc) goto done; }
        }
        te.cost = x_cost; te.radius = radius; te.mu = mu; te.gradient_max_norm = gradient_max_norm;
        trace_push(trace, &te); sum->num_iterations++;
        /* FinalizeIterationAndCheckIfMinimizerCanContinue */
        if (iteration >= opt->max_num_iterations) { term = VIWB_NO_CONVERGENCE; break; }
        if (te.step_is_successful && gradient_max_norm <= opt->gradient_tolerance) { term = VIWB_CONVERGENCE; break; }
        if (radius <= opt->min_trust_region_radius) { term = VIWB_CONVERGENCE; break; }
    }
done:
    if (!rc) {
        memcpy(state, x, sizeof(double) * nstate);
        sum->termination_type = term; sum->num_successful_steps = successful_steps; sum->num_linear_solves = num_linear;
        sum->final_cost = x_cost + P.fixed_cost; sum->final_radius = radius; sum->final_mu = mu;
    }
    schur_free(&S); free(tmpm); free(scale); free(x); program_free(&P);
    return rc;
#undef EVAL_GRADIENT_AND_JACOBIAN
}

/* ------------------------------------------------------------------ debug / cross-check hooks */
int vo_cost(const viwb_problem *pb, const double *state, double *cost, double *residuals) {
    program_t P; int rc = program_build(&P, pb);
    if (rc) { program_free(&P); return rc; }
    double c; rc = program_evaluate(&P, state, 0, &c);
    if (rc) { program_free(&P); return rc; }
    *cost = c + P.fixed_cost;
    if (residuals) memcpy(residuals, P.residuals, sizeof(double) * P.nrows);
    int rows = P.nrows; program_free(&P);
    return rows;
}

int vo_normal_equations(const viwb_problem *pb, const double *state, double *H, double *g, double *lm, double *cost) {
    program_t P; int rc = program_build(&P, pb);
    if (rc) { program_free(&P); return rc; }
    double c; rc = program_evaluate(&P, state, 1, &c);
    if (rc) { program_free(&P); return rc; }
    *cost = c + P.fixed_cost;
    const int T = VIWB_TANGENT_FIXED;
    memset(H, 0, sizeof(double) * T * T); memset(g, 0, sizeof(double) * T);
    if (lm) memset(lm, 0, sizeof(double) * 82 * pb->num_landmarks);
    for (int r = 0; r < P.nrb; r++) {
        const rblock_t *R = &P.rb[r]; const double *res = P.residuals + R->row0;
        for (int a = 0; a < R->nslots; a++) {
            if (R->col[a] < 0) continue;
            int ba = R->block[a], ta = R->tsize[a]; const double *Ja = P.jac + R->jac0[a];
            for (int i = 0; i < R->nrows; i++) for (int cidx = 0; cidx < ta; cidx++) {
                double v = Ja[i * ta + cidx] * res[i];
                if (ba < VIWB_NUM_FIXED_BLOCKS) g[viwb_block_toffset(ba) + cidx] += v; else if (lm) lm[(ba - 32) * 82 + 1] += v;
            }
            for (int b = 0; b < R->nslots; b++) {
                if (R->col[b] < 0) continue;
                int bb = R->block[b], tb = R->tsize[b]; const double *Jb = P.jac + R->jac0[b];
                for (int cidx = 0; cidx < ta; cidx++) for (int d = 0; d < tb; d++) {
                    double v = 0; for (int i = 0; i < R->nrows; i++) v += Ja[i * ta + cidx] * Jb[i * tb + d];
                    if (ba < 32 && bb < 32) H[(viwb_block_toffset(ba) + cidx) * T + viwb_block_toffset(bb) + d] += v;
                    else if (ba >= 32 && bb >= 32) { if (lm) lm[(ba - 32) * 82] += v; }
                    else if (ba >= 32 && lm) {
                        /* w_k over the visual subspace: poses 0..65, ex0 66..71, ex1 72..77, td 78 */
                        int vs = bb <= 10 ? 6 * bb : bb == VIWB_BLK_EX_POSE0 ? 66 : bb == VIWB_BLK_EX_POSE1 ? 72 : bb == VIWB_BLK_TD ? 78 : -1;
                        if (vs >= 0) lm[(ba - 32) * 82 + 2 + vs + d] += v;
                    }
                }
            }
        }
    }
    program_free(&P);
    return 0;
}

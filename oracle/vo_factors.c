/*
 * vo_factors.c -- CPU oracle, part 1: literal FP64 restatements of the reference's factor classes.
 * TEST INFRASTRUCTURE ONLY (see viw_oracle.h). Paths are relative to /root/reference/vins_estimator/src.
 */
#include "viw_oracle.h"
#include "vo_math.h"
#include <stdlib.h>

/* ---------------------------------------------------------------- small dense helpers */
/* inverse by LU with partial pivoting (what Eigen's MatrixXd/fixed >4 ::inverse() does: PartialPivLU) */
static int dense_inverse(int n, const double *A, double *Ainv) {
    double *a = (double *)malloc(sizeof(double) * n * n * 2);
    double *inv = a + n * n;
    memcpy(a, A, sizeof(double) * n * n);
    for (int i = 0; i < n * n; i++) inv[i] = 0;
    for (int i = 0; i < n; i++) inv[i * n + i] = 1;
    for (int c = 0; c < n; c++) {
        int p = c; double best = fabs(a[c * n + c]);
        for (int r = c + 1; r < n; r++) if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); p = r; }
        if (best == 0) { free(a); return -1; }
        if (p != c) for (int k = 0; k < n; k++) {
            double t = a[c * n + k]; a[c * n + k] = a[p * n + k]; a[p * n + k] = t;
            t = inv[c * n + k]; inv[c * n + k] = inv[p * n + k]; inv[p * n + k] = t;
        }
        double d = a[c * n + c];
        for (int r = c + 1; r < n; r++) {
            double f = a[r * n + c] / d;
            if (f == 0) continue;
            for (int k = c; k < n; k++) a[r * n + k] -= f * a[c * n + k];
            for (int k = 0; k < n; k++) inv[r * n + k] -= f * inv[c * n + k];
        }
    }
    for (int c = n - 1; c >= 0; c--) {
        double d = a[c * n + c];
        for (int k = 0; k < n; k++) inv[c * n + k] /= d;
        for (int r = 0; r < c; r++) {
            double f = a[r * n + c];
            if (f == 0) continue;
            for (int k = 0; k < n; k++) inv[r * n + k] -= f * inv[c * n + k];
        }
    }
    memcpy(Ainv, inv, sizeof(double) * n * n);
    free(a);
    return 0;
}
/* Eigen::LLT lower factor; returns -1 on a non-positive pivot */
static int dense_llt(int n, const double *A, double *L) {
    for (int i = 0; i < n * n; i++) L[i] = 0;
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0)) return -1;
        d = sqrt(d); L[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / d;
        }
    }
    return 0;
}
/* sqrt_info = LLT(cov.inverse()).matrixL().transpose()   (imu_factor.h:75, wheel_factor.h:85) */
static int sqrt_info_from_cov(int n, const double *cov, double *S) {
    double inv[225], L[225];
    if (dense_inverse(n, cov, inv)) return -1;
    if (dense_llt(n, inv, L)) return -1;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) S[i * n + j] = L[j * n + i];
    return 0;
}
/* J (rows x cols, row-major) <- S (rows x rows) * J */
static void left_mul_inplace(int rows, int cols, const double *S, double *J) {
    double *t = (double *)malloc(sizeof(double) * rows * cols);
    mat_mul(t, S, J, rows, rows, cols);
    memcpy(J, t, sizeof(double) * rows * cols);
    free(t);
}
/* write a 3x3 block into a row-major matrix with `ld` columns */
static void put33(double *J, int ld, int r0, int c0, const double *m) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[(r0 + i) * ld + c0 + j] = m[i * 3 + j];
}
static void q_left44(double *o, const double *q) { /* Utility::Qleft (utility.h:58-66), w first */
    double s[9]; m3_skew(s, q);
    o[0] = q[3]; o[1] = -q[0]; o[2] = -q[1]; o[3] = -q[2];
    for (int i = 0; i < 3; i++) { o[(i + 1) * 4] = q[i]; for (int j = 0; j < 3; j++) o[(i + 1) * 4 + 1 + j] = (i == j ? q[3] : 0) + s[i * 3 + j]; }
}
static void q_right44(double *o, const double *q) { /* Utility::Qright (utility.h:68-76) */
    double s[9]; m3_skew(s, q);
    o[0] = q[3]; o[1] = -q[0]; o[2] = -q[1]; o[3] = -q[2];
    for (int i = 0; i < 3; i++) { o[(i + 1) * 4] = q[i]; for (int j = 0; j < 3; j++) o[(i + 1) * 4 + 1 + j] = (i == j ? q[3] : 0) - s[i * 3 + j]; }
}

/* ================================================================ visual factors */
/* shared head of the three projection factors; obs = pts_i(3) pts_j(3) vel_i(2) vel_j(2) td_i td_j */
typedef struct { double pts_i[3], pts_j[3], vel_i[3], vel_j[3], td_i, td_j; } vis_obs_t;
static void load_obs(vis_obs_t *o, const double *c) {
    v3_copy(o->pts_i, c); v3_copy(o->pts_j, c + 3);
    o->vel_i[0] = c[6]; o->vel_i[1] = c[7]; o->vel_i[2] = 0;
    o->vel_j[0] = c[8]; o->vel_j[1] = c[9]; o->vel_j[2] = 0;
    o->td_i = c[10]; o->td_j = c[11];
}
/* reduce (2x3) = sqrt_info * [1/z 0 -x/z^2; 0 1/z -y/z^2] */
static void make_reduce(double *reduce, const double *S, const double *pc) {
    double dep = pc[2];
    double r[6] = {1. / dep, 0, -pc[0] / (dep * dep), 0, 1. / dep, -pc[1] / (dep * dep)};
    mat_mul(reduce, S, r, 2, 2, 3);
}
/* jac (2x7 row-major) <- [reduce * jaco(3x6), 0] */
static void put_pose_jac(double *jac, const double *reduce, const double *jaco /*3x6*/) {
    double t[12]; mat_mul(t, reduce, jaco, 2, 3, 6);
    for (int r = 0; r < 2; r++) { for (int c = 0; c < 6; c++) jac[r * 7 + c] = t[r * 6 + c]; jac[r * 7 + 6] = 0; }
}
static void cat36(double *o, const double *L, const double *R) { /* [L | R], 3x3 each -> 3x6 */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { o[i * 6 + j] = L[i * 3 + j]; o[i * 6 + 3 + j] = R[i * 3 + j]; }
}

/* ProjectionTwoFrameOneCamFactor::Evaluate (factor/projectionTwoFrameOneCamFactor.cpp:45-152) */
static int eval_proj_2f1c(const viwb_globals *g, const double *consts, const double *const *p, double *res, double **jac) {
    vis_obs_t o; load_obs(&o, consts);
    const double *Pi = p[0], *Qi = p[0] + 3, *Pj = p[1], *Qj = p[1] + 3, *tic = p[2], *qic = p[2] + 3;
    double inv_dep_i = p[3][0], td = p[4][0];
    double pts_i_td[3], pts_j_td[3], t[3];
    v3_scale(t, o.vel_i, td - o.td_i); v3_sub(pts_i_td, o.pts_i, t);
    v3_scale(t, o.vel_j, td - o.td_j); v3_sub(pts_j_td, o.pts_j, t);
    double pts_camera_i[3], pts_imu_i[3], pts_w[3], pts_imu_j[3], pts_camera_j[3], qinv[4];
    v3_scale(pts_camera_i, pts_i_td, 1.0 / inv_dep_i);
    q_rot(t, qic, pts_camera_i); v3_add(pts_imu_i, t, tic);
    q_rot(t, Qi, pts_imu_i); v3_add(pts_w, t, Pi);
    v3_sub(t, pts_w, Pj); q_inv(qinv, Qj); q_rot(pts_imu_j, qinv, t);
    v3_sub(t, pts_imu_j, tic); q_inv(qinv, qic); q_rot(pts_camera_j, qinv, t);
    double dep_j = pts_camera_j[2];
    double r0[2] = {pts_camera_j[0] / dep_j - pts_j_td[0], pts_camera_j[1] / dep_j - pts_j_td[1]};
    const double *S = g->vis_sqrt_info;
    res[0] = S[0] * r0[0] + S[1] * r0[1]; res[1] = S[2] * r0[0] + S[3] * r0[1];
    if (!jac) return 0;
    double Ri[9], Rj[9], ric[9], reduce[6];
    q_to_R(Ri, Qi); q_to_R(Rj, Qj); q_to_R(ric, qic);
    make_reduce(reduce, S, pts_camera_j);
    double ricT[9], RjT[9], ricT_RjT[9], sk[9], L[9], Rr[9], jaco[18];
    m3_transpose(ricT, ric); m3_transpose(RjT, Rj); m3_mul(ricT_RjT, ricT, RjT);
    if (jac[0]) {
        m3_skew(sk, pts_imu_i); m3_scale(sk, sk, -1.0);
        m3_mul(Rr, ricT_RjT, Ri); m3_mul(Rr, Rr, sk);
        cat36(jaco, ricT_RjT, Rr); put_pose_jac(jac[0], reduce, jaco);
    }
    if (jac[1]) {
        m3_scale(L, ricT_RjT, -1.0);
        m3_skew(sk, pts_imu_j); m3_mul(Rr, ricT, sk);
        cat36(jaco, L, Rr); put_pose_jac(jac[1], reduce, jaco);
    }
    double tmp_r[9]; m3_mul(tmp_r, ricT_RjT, Ri); m3_mul(tmp_r, tmp_r, ric);
    if (jac[2]) {
        double RjT_Ri[9], I[9]; m3_mul(RjT_Ri, RjT, Ri); m3_identity(I); m3_sub(RjT_Ri, RjT_Ri, I);
        m3_mul(L, ricT, RjT_Ri);
        double a[9], b[9], c[9], v[3], w[3];
        m3_skew(sk, pts_camera_i); m3_mul(a, tmp_r, sk); m3_scale(a, a, -1.0);
        m3_mulv(v, tmp_r, pts_camera_i); m3_skew(b, v);
        /* ric^T (Rj^T (Ri tic + Pi - Pj) - tic) */
        m3_mulv(v, Ri, tic); v3_add(v, v, Pi); v3_sub(v, v, Pj); m3_mulv(w, RjT, v); v3_sub(w, w, tic); m3_mulv(v, ricT, w);
        m3_skew(c, v);
        m3_add(Rr, a, b); m3_add(Rr, Rr, c);
        cat36(jaco, L, Rr); put_pose_jac(jac[2], reduce, jaco);
    }
    if (jac[3]) {
        double v[3], w[2];
        m3_mulv(v, tmp_r, pts_i_td); mat_mul(w, reduce, v, 2, 3, 1);
        jac[3][0] = w[0] * -1.0 / (inv_dep_i * inv_dep_i); jac[3][1] = w[1] * -1.0 / (inv_dep_i * inv_dep_i);
    }
    if (jac[4]) {
        double v[3], w[2];
        m3_mulv(v, tmp_r, o.vel_i); mat_mul(w, reduce, v, 2, 3, 1);
        jac[4][0] = w[0] / inv_dep_i * -1.0 + (S[0] * o.vel_j[0] + S[1] * o.vel_j[1]);
        jac[4][1] = w[1] / inv_dep_i * -1.0 + (S[2] * o.vel_j[0] + S[3] * o.vel_j[1]);
    }
    return 0;
}

/* ProjectionTwoFrameTwoCamFactor::Evaluate (factor/projectionTwoFrameTwoCamFactor.cpp:43-166) */
static int eval_proj_2f2c(const viwb_globals *g, const double *consts, const double *const *p, double *res, double **jac) {
    vis_obs_t o; load_obs(&o, consts);
    const double *Pi = p[0], *Qi = p[0] + 3, *Pj = p[1], *Qj = p[1] + 3, *tic = p[2], *qic = p[2] + 3, *tic2 = p[3], *qic2 = p[3] + 3;
    double inv_dep_i = p[4][0], td = p[5][0];
    double pts_i_td[3], pts_j_td[3], t[3];
    v3_scale(t, o.vel_i, td - o.td_i); v3_sub(pts_i_td, o.pts_i, t);
    v3_scale(t, o.vel_j, td - o.td_j); v3_sub(pts_j_td, o.pts_j, t);
    double pts_camera_i[3], pts_imu_i[3], pts_w[3], pts_imu_j[3], pts_camera_j[3], qinv[4];
    v3_scale(pts_camera_i, pts_i_td, 1.0 / inv_dep_i);
    q_rot(t, qic, pts_camera_i); v3_add(pts_imu_i, t, tic);
    q_rot(t, Qi, pts_imu_i); v3_add(pts_w, t, Pi);
    v3_sub(t, pts_w, Pj); q_inv(qinv, Qj); q_rot(pts_imu_j, qinv, t);
    v3_sub(t, pts_imu_j, tic2); q_inv(qinv, qic2); q_rot(pts_camera_j, qinv, t);
    double dep_j = pts_camera_j[2];
    double r0[2] = {pts_camera_j[0] / dep_j - pts_j_td[0], pts_camera_j[1] / dep_j - pts_j_td[1]};
    const double *S = g->vis_sqrt_info;
    res[0] = S[0] * r0[0] + S[1] * r0[1]; res[1] = S[2] * r0[0] + S[3] * r0[1];
    if (!jac) return 0;
    double Ri[9], Rj[9], ric[9], ric2[9], reduce[6];
    q_to_R(Ri, Qi); q_to_R(Rj, Qj); q_to_R(ric, qic); q_to_R(ric2, qic2);
    make_reduce(reduce, S, pts_camera_j);
    double ric2T[9], RjT[9], A[9], sk[9], L[9], Rr[9], jaco[18];
    m3_transpose(ric2T, ric2); m3_transpose(RjT, Rj); m3_mul(A, ric2T, RjT); /* A = ric2^T Rj^T */
    if (jac[0]) {
        m3_skew(sk, pts_imu_i); m3_scale(sk, sk, -1.0);
        m3_mul(Rr, A, Ri); m3_mul(Rr, Rr, sk);
        cat36(jaco, A, Rr); put_pose_jac(jac[0], reduce, jaco);
    }
    if (jac[1]) {
        m3_scale(L, A, -1.0); m3_skew(sk, pts_imu_j); m3_mul(Rr, ric2T, sk);
        cat36(jaco, L, Rr); put_pose_jac(jac[1], reduce, jaco);
    }
    double ARi[9], ARiric[9];
    m3_mul(ARi, A, Ri); m3_mul(ARiric, ARi, ric);
    if (jac[2]) {
        m3_skew(sk, pts_camera_i); m3_scale(sk, sk, -1.0); m3_mul(Rr, ARiric, sk);
        cat36(jaco, ARi, Rr); put_pose_jac(jac[2], reduce, jaco);
    }
    if (jac[3]) {
        m3_scale(L, ric2T, -1.0); m3_skew(Rr, pts_camera_j);
        cat36(jaco, L, Rr); put_pose_jac(jac[3], reduce, jaco);
    }
    if (jac[4]) {
        double v[3], w[2];
        m3_mulv(v, ARiric, pts_i_td); mat_mul(w, reduce, v, 2, 3, 1);
        jac[4][0] = w[0] * -1.0 / (inv_dep_i * inv_dep_i); jac[4][1] = w[1] * -1.0 / (inv_dep_i * inv_dep_i);
    }
    if (jac[5]) {
        double v[3], w[2];
        m3_mulv(v, ARiric, o.vel_i); mat_mul(w, reduce, v, 2, 3, 1);
        jac[5][0] = w[0] / inv_dep_i * -1.0 + (S[0] * o.vel_j[0] + S[1] * o.vel_j[1]);
        jac[5][1] = w[1] / inv_dep_i * -1.0 + (S[2] * o.vel_j[0] + S[3] * o.vel_j[1]);
    }
    return 0;
}

/* ProjectionOneFrameTwoCamFactor::Evaluate (factor/projectionOneFrameTwoCamFactor.cpp:42-134) */
static int eval_proj_1f2c(const viwb_globals *g, const double *consts, const double *const *p, double *res, double **jac) {
    vis_obs_t o; load_obs(&o, consts);
    const double *tic = p[0], *qic = p[0] + 3, *tic2 = p[1], *qic2 = p[1] + 3;
    double inv_dep_i = p[2][0], td = p[3][0];
    double pts_i_td[3], pts_j_td[3], t[3];
    v3_scale(t, o.vel_i, td - o.td_i); v3_sub(pts_i_td, o.pts_i, t);
    v3_scale(t, o.vel_j, td - o.td_j); v3_sub(pts_j_td, o.pts_j, t);
    double pts_camera_i[3], pts_imu_i[3], pts_camera_j[3], qinv[4];
    v3_scale(pts_camera_i, pts_i_td, 1.0 / inv_dep_i);
    q_rot(t, qic, pts_camera_i); v3_add(pts_imu_i, t, tic);
    v3_sub(t, pts_imu_i, tic2); q_inv(qinv, qic2); q_rot(pts_camera_j, qinv, t);
    double dep_j = pts_camera_j[2];
    double r0[2] = {pts_camera_j[0] / dep_j - pts_j_td[0], pts_camera_j[1] / dep_j - pts_j_td[1]};
    const double *S = g->vis_sqrt_info;
    res[0] = S[0] * r0[0] + S[1] * r0[1]; res[1] = S[2] * r0[0] + S[3] * r0[1];
    if (!jac) return 0;
    double ric[9], ric2[9], reduce[6], ric2T[9], B[9], sk[9], L[9], Rr[9], jaco[18];
    q_to_R(ric, qic); q_to_R(ric2, qic2);
    make_reduce(reduce, S, pts_camera_j);
    m3_transpose(ric2T, ric2); m3_mul(B, ric2T, ric); /* B = ric2^T ric */
    if (jac[0]) {
        m3_skew(sk, pts_camera_i); m3_scale(sk, sk, -1.0); m3_mul(Rr, B, sk);
        cat36(jaco, ric2T, Rr); put_pose_jac(jac[0], reduce, jaco);
    }
    if (jac[1]) {
        m3_scale(L, ric2T, -1.0); m3_skew(Rr, pts_camera_j);
        cat36(jaco, L, Rr); put_pose_jac(jac[1], reduce, jaco);
    }
    if (jac[2]) { /* quirk: pts_i, not pts_i_td (projectionOneFrameTwoCamFactor.cpp:119) */
        double v[3], w[2];
        m3_mulv(v, B, o.pts_i); mat_mul(w, reduce, v, 2, 3, 1);
        jac[2][0] = w[0] * -1.0 / (inv_dep_i * inv_dep_i); jac[2][1] = w[1] * -1.0 / (inv_dep_i * inv_dep_i);
    }
    if (jac[3]) {
        double v[3], w[2];
        m3_mulv(v, B, o.vel_i); mat_mul(w, reduce, v, 2, 3, 1);
        jac[3][0] = w[0] / inv_dep_i * -1.0 + (S[0] * o.vel_j[0] + S[1] * o.vel_j[1]);
        jac[3][1] = w[1] / inv_dep_i * -1.0 + (S[2] * o.vel_j[0] + S[3] * o.vel_j[1]);
    }
    return 0;
}

/* ================================================================ IMU factor */
/* IMUFactor::Evaluate (factor/imu_factor.h:30-192) + IntegrationBase::evaluate (integration_base.h:169-195) */
static int eval_imu(const viwb_globals *g, const double *c, const double *const *p, double *res, double **jac) {
    const double sum_dt = c[0], *delta_p = c + 1, *delta_q = c + 4, *delta_v = c + 8, *lin_ba = c + 11, *lin_bg = c + 14;
    const double *dp_dba = c + 17, *dp_dbg = c + 26, *dq_dbg = c + 35, *dv_dba = c + 44, *dv_dbg = c + 53, *cov = c + 62;
    const double *Pi = p[0], *Qi = p[0] + 3, *Vi = p[1], *Bai = p[1] + 3, *Bgi = p[1] + 6;
    const double *Pj = p[2], *Qj = p[2] + 3, *Vj = p[3], *Baj = p[3] + 3, *Bgj = p[3] + 6;
    const double *G = g->G;
    double dba[3], dbg[3], t[3], u[3], corrected_delta_q[4], dq[4], corrected_delta_v[3], corrected_delta_p[3];
    v3_sub(dba, Bai, lin_ba); v3_sub(dbg, Bgi, lin_bg);
    m3_mulv(t, dq_dbg, dbg); q_delta(dq, t); q_mul(corrected_delta_q, delta_q, dq);
    m3_mulv(t, dv_dba, dba); m3_mulv(u, dv_dbg, dbg); v3_add(corrected_delta_v, delta_v, t); v3_add(corrected_delta_v, corrected_delta_v, u);
    m3_mulv(t, dp_dba, dba); m3_mulv(u, dp_dbg, dbg); v3_add(corrected_delta_p, delta_p, t); v3_add(corrected_delta_p, corrected_delta_p, u);
    double Qi_inv[4], a[3], r[15];
    q_inv(Qi_inv, Qi);
    /* r_p = Qi^-1 (0.5 G dt^2 + Pj - Pi - Vi dt) - corrected_delta_p */
    for (int k = 0; k < 3; k++) a[k] = 0.5 * G[k] * sum_dt * sum_dt + Pj[k] - Pi[k] - Vi[k] * sum_dt;
    q_rot(t, Qi_inv, a); v3_sub(r + 0, t, corrected_delta_p);
    /* r_q = 2 (corrected_delta_q^-1 (Qi^-1 Qj)).vec */
    double qij[4], cq_inv[4], qe[4];
    q_mul(qij, Qi_inv, Qj); q_inv(cq_inv, corrected_delta_q); q_mul(qe, cq_inv, qij);
    r[3] = 2 * qe[0]; r[4] = 2 * qe[1]; r[5] = 2 * qe[2];
    for (int k = 0; k < 3; k++) a[k] = G[k] * sum_dt + Vj[k] - Vi[k];
    q_rot(t, Qi_inv, a); v3_sub(r + 6, t, corrected_delta_v);
    v3_sub(r + 9, Baj, Bai); v3_sub(r + 12, Bgj, Bgi);
    double S[225];
    if (sqrt_info_from_cov(15, cov, S)) return -1;
    mat_mul(res, S, r, 15, 15, 1);
    if (!jac) return 0;
    double RiT[9], sk[9], m[9], Qj_inv[4];
    q_to_R(m, Qi_inv); m3_copy(RiT, m); /* Qi.inverse().toRotationMatrix() */
    q_inv(Qj_inv, Qj);
    if (jac[0]) {
        double *J = jac[0]; memset(J, 0, sizeof(double) * 15 * 7);
        m3_scale(m, RiT, -1.0); put33(J, 7, 0, 0, m);
        for (int k = 0; k < 3; k++) a[k] = 0.5 * G[k] * sum_dt * sum_dt + Pj[k] - Pi[k] - Vi[k] * sum_dt;
        q_rot(t, Qi_inv, a); m3_skew(sk, t); put33(J, 7, 0, 3, sk);
        /* -(Qleft(Qj^-1 Qi) * Qright(corrected_delta_q)).bottomRightCorner<3,3>() */
        double qji[4], L4[16], R4[16], LR[16];
        q_mul(qji, Qj_inv, Qi); q_left44(L4, qji); q_right44(R4, corrected_delta_q); mat_mul(LR, L4, R4, 4, 4, 4);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i * 3 + j] = -LR[(i + 1) * 4 + j + 1];
        put33(J, 7, 3, 3, m);
        for (int k = 0; k < 3; k++) a[k] = G[k] * sum_dt + Vj[k] - Vi[k];
        q_rot(t, Qi_inv, a); m3_skew(sk, t); put33(J, 7, 6, 3, sk);
        left_mul_inplace(15, 7, S, J);
    }
    if (jac[1]) {
        double *J = jac[1]; memset(J, 0, sizeof(double) * 15 * 9);
        m3_scale(m, RiT, -sum_dt); put33(J, 9, 0, 0, m);
        m3_scale(m, dp_dba, -1.0); put33(J, 9, 0, 3, m);
        m3_scale(m, dp_dbg, -1.0); put33(J, 9, 0, 6, m);
        /* -Qleft(Qj^-1 Qi delta_q).bottomRightCorner<3,3>() * dq_dbg   (quirk 5: un-corrected delta_q, imu_factor.h:138) */
        double q1[4], q2[4], l33[9];
        q_mul(q1, Qj_inv, Qi); q_mul(q2, q1, delta_q); q_left33(l33, q2); m3_mul(m, l33, dq_dbg); m3_scale(m, m, -1.0);
        put33(J, 9, 3, 6, m);
        m3_scale(m, RiT, -1.0); put33(J, 9, 6, 0, m);
        m3_scale(m, dv_dba, -1.0); put33(J, 9, 6, 3, m);
        m3_scale(m, dv_dbg, -1.0); put33(J, 9, 6, 6, m);
        m3_identity(m); m3_scale(m, m, -1.0); put33(J, 9, 9, 3, m); put33(J, 9, 12, 6, m);
        left_mul_inplace(15, 9, S, J);
    }
    if (jac[2]) {
        double *J = jac[2]; memset(J, 0, sizeof(double) * 15 * 7);
        put33(J, 7, 0, 0, RiT);
        double q1[4], q2[4], l33[9];
        q_mul(q1, cq_inv, Qi_inv); q_mul(q2, q1, Qj); q_left33(l33, q2);
        put33(J, 7, 3, 3, l33);
        left_mul_inplace(15, 7, S, J);
    }
    if (jac[3]) {
        double *J = jac[3]; memset(J, 0, sizeof(double) * 15 * 9);
        put33(J, 9, 6, 0, RiT);
        m3_identity(m); put33(J, 9, 9, 3, m); put33(J, 9, 12, 6, m);
        left_mul_inplace(15, 9, S, J);
    }
    return 0;
}

/* ================================================================ wheel factor */
/* WheelFactor::Evaluate (factor/wheel_factor.h:28-246) + WheelIntegrationBase::evaluate (wheel_integration_base.h:179-218) */
static int eval_wheel(const viwb_globals *g, const double *c, const double *const *p, double *res, double **jac) {
    (void)g;
    const double *delta_p = c, *delta_q = c + 3, *Jpre = c + 7, *cov = c + 25;
    const double lin_sx = c[61], lin_sy = c[62], lin_sw = c[63], lin_td = c[64];
    const double *lin_vel = c + 65, *lin_gyr = c + 68, *vel_1 = c + 71, *gyr_1 = c + 74;
    const double *Pi = p[0], *Qi = p[0] + 3, *Pj = p[1], *Qj = p[1] + 3, *tio = p[2], *qio = p[2] + 3;
    const double sx = p[3][0], sy = p[4][0], sw = p[5][0], td = p[6][0];
    double dp_dsx[3] = {Jpre[0], Jpre[3], Jpre[6]}, dp_dsy[3] = {Jpre[1], Jpre[4], Jpre[7]}, dp_dsw[3] = {Jpre[2], Jpre[5], Jpre[8]};
    double dq_dsw[3] = {Jpre[11], Jpre[14], Jpre[17]};
    double dsx = sx - lin_sx, dsy = sy - lin_sy, dsw = sw - lin_sw;
    double sv[3] = {sx, sy, 1.0};
    double Ri[9], Rj[9], rio[9];
    q_to_R(Ri, Qi); q_to_R(Rj, Qj); q_to_R(rio, qio);
    /* ---- evaluate() ---- */
    double corrected_delta_p[3], corrected_delta_q[4], t[3], u[3], w[3], qa[4], qb[4], qn[4];
    for (int k = 0; k < 3; k++) corrected_delta_p[k] = delta_p[k] + dp_dsx[k] * dsx + dp_dsy[k] * dsy + dp_dsw[k] * dsw;
    v3_scale(t, dq_dsw, dsw); so3_exp_q(qa, t);
    memcpy(qn, delta_q, sizeof qn); q_normalize(qn);          /* Sophus::SO3d(delta_q) normalises */
    q_mul(corrected_delta_q, qn, qa);
    double dtd = td - lin_td;
    double fw[3], bw[3], qfw[4], qbw[4], delta_q_time[4], Rfw[9];
    v3_scale(fw, lin_gyr, sw * dtd); v3_scale(bw, gyr_1, -sw * dtd);
    so3_exp_q(qfw, fw); so3_exp_q(qbw, bw);
    memcpy(qn, corrected_delta_q, sizeof qn); q_normalize(qn);
    q_mul(qb, qfw, qn); q_mul(delta_q_time, qb, qbw);
    q_to_R(Rfw, qfw);
    double sv_lv[3] = {sv[0] * lin_vel[0], sv[1] * lin_vel[1], sv[2] * lin_vel[2]};
    double sv_v1[3] = {sv[0] * vel_1[0], sv[1] * vel_1[1], sv[2] * vel_1[2]};
    v3_scale(t, sv_v1, dtd); q_rot(u, corrected_delta_q, t);              /* corrected_delta_q * sv * vel_1 * dtd */
    for (int k = 0; k < 3; k++) w[k] = sv_lv[k] * dtd + corrected_delta_p[k] - u[k];
    double delta_p_time[3]; m3_mulv(delta_p_time, Rfw, w);
    double Rio_w[9], d[3], raw[6];
    m3_mul(Rio_w, Ri, rio);                                               /* Ri * rio */
    m3_mulv(t, Rj, tio); m3_mulv(u, Ri, tio);
    for (int k = 0; k < 3; k++) d[k] = t[k] + Pj[k] - u[k] - Pi[k];      /* Rj tio + Pj - Ri tio - Pi */
    m3_tmulv(t, Rio_w, d); v3_sub(raw, t, delta_p_time);
    double q_iio[4], q_iio_inv[4], dqt_inv[4], q1[4], q2[4], q3[4];
    q_mul(q_iio, Qi, qio); q_inv(q_iio_inv, q_iio); q_inv(dqt_inv, delta_q_time);
    q_mul(q1, dqt_inv, q_iio_inv); q_mul(q2, q1, Qj); q_mul(q3, q2, qio);
    so3_log_q(raw + 3, q3);
    double S[36];
    if (sqrt_info_from_cov(6, cov, S)) return -1;
    mat_mul(res, S, raw, 6, 6, 1);
    if (!jac) return 0;
    /* ---- Jacobians ---- */
    double *raw_r = raw + 3, Jr_dq_inv[9], drdsw[3], Jr_drdsw[9], m[9], n[9], sk[9];
    so3_Jr_inv(Jr_dq_inv, raw_r);
    v3_scale(drdsw, dq_dsw, sw - lin_sw); so3_Jr(Jr_drdsw, drdsw);
    double R_iio_inv[9]; q_to_R(R_iio_inv, q_iio_inv);                    /* (Qi*qio).inverse().toRotationMatrix() */
    if (jac[0]) {
        double *J = jac[0]; memset(J, 0, sizeof(double) * 42);
        m3_scale(m, R_iio_inv, -1.0); put33(J, 7, 0, 0, m);
        /* (ri rio)^T (ri skew(tio)) + rio^T skew(ri^T (rj tio + Pj - ri tio - Pi)) */
        double RioT[9], a[9], b[9];
        m3_transpose(RioT, Rio_w); m3_skew(sk, tio); m3_mul(a, Ri, sk); m3_mul(a, RioT, a);
        m3_tmulv(t, Ri, d); m3_skew(sk, t); m3_transpose(n, rio); m3_mul(b, n, sk);
        m3_add(m, a, b); put33(J, 7, 0, 3, m);
        /* -Jr_inv * ((Qj*qio).inverse() * Qi).toRotationMatrix() */
        q_mul(q1, Qj, qio); q_inv(q2, q1); q_mul(q3, q2, Qi); q_to_R(n, q3);
        m3_mul(m, Jr_dq_inv, n); m3_scale(m, m, -1.0); put33(J, 7, 3, 3, m);
        left_mul_inplace(6, 7, S, J);
    }
    if (jac[1]) {
        double *J = jac[1]; memset(J, 0, sizeof(double) * 42);
        put33(J, 7, 0, 0, R_iio_inv);
        q_mul(q1, q_iio_inv, Qj); q_to_R(n, q1); m3_skew(sk, tio); m3_mul(m, n, sk); m3_scale(m, m, -1.0);
        put33(J, 7, 0, 3, m);
        q_inv(q1, qio); q_to_R(n, q1); m3_mul(m, Jr_dq_inv, n); put33(J, 7, 3, 3, m);
        left_mul_inplace(6, 7, S, J);
    }
    if (jac[2]) {
        double *J = jac[2]; memset(J, 0, sizeof(double) * 42);
        m3_sub(n, Rj, Ri); m3_mul(m, R_iio_inv, n); put33(J, 7, 0, 0, m);
        /* skew((Qi*qio).inverse() * (Qj*tio + Pj - Qi*tio - Pi)) */
        double e[3]; q_rot(t, Qj, tio); q_rot(u, Qi, tio);
        for (int k = 0; k < 3; k++) e[k] = t[k] + Pj[k] - u[k] - Pi[k];
        q_rot(t, q_iio_inv, e); m3_skew(sk, t); put33(J, 7, 0, 3, sk);
        /* Jr_inv * (I - ((Qj*qio).inverse() * Qi * qio).toRotationMatrix()) */
        q_mul(q1, Qj, qio); q_inv(q2, q1); q_mul(q3, q2, Qi); q_mul(q1, q3, qio); q_to_R(n, q1);
        m3_identity(m); m3_sub(m, m, n); m3_mul(n, Jr_dq_inv, m); put33(J, 7, 3, 3, n);
        left_mul_inplace(6, 7, S, J);
    }
    double fcw[3], fcv[3], bcv[3], bcw[3], Jrtd[9], Jr_minus_td[9], nfw[3];
    v3_scale(fcw, lin_gyr, sw * dtd);
    for (int k = 0; k < 3; k++) { fcv[k] = sv[k] * lin_vel[k] * dtd; bcv[k] = sv[k] * vel_1[k] * dtd; }
    v3_scale(bcw, gyr_1, sw * dtd);
    so3_Jr(Jrtd, fcw); v3_scale(nfw, fcw, -1.0); so3_Jr(Jr_minus_td, nfw);
    double Rcdq[9]; q_to_R(Rcdq, corrected_delta_q);                     /* corrected_delta_q.toRotationMatrix() */
    double Efv[9], Efw[9];
    so3_exp_R(Efv, fcv); so3_exp_R(Efw, fcw);
    for (int axis = 0; axis < 2; axis++) {                               /* sx (I1) and sy (I2); quirk 4: Exp(forward_compensate_v) */
        double *J = jac[3 + axis]; if (!J) continue;
        const double *dp_ds = axis == 0 ? dp_dsx : dp_dsy;
        double Ilv[3] = {0, 0, 0}, Iv1[3] = {0, 0, 0}, v[3];
        Ilv[axis] = lin_vel[axis] * dtd; Iv1[axis] = vel_1[axis] * dtd;
        m3_mulv(t, Rcdq, Iv1);
        for (int k = 0; k < 3; k++) v[k] = Ilv[k] + dp_ds[k] - t[k];
        m3_mulv(t, Efv, v);
        double col[6] = {-t[0], -t[1], -t[2], 0, 0, 0};
        mat_mul(J, S, col, 6, 6, 1);
    }
    if (jac[5]) {
        double v[3], a[3], b[3], col[6];
        /* dp_dsw - Rcdq skew(Jr_drdsw dq_dsw) sv vel_1 dtd + skew(Jrtd lin_gyr dtd) (fcv + cdp - cdq * bcv) */
        m3_mulv(t, Jr_drdsw, dq_dsw); m3_skew(sk, t); v3_scale(u, sv_v1, dtd); m3_mulv(a, sk, u); m3_mulv(a, Rcdq, a);
        v3_scale(u, lin_gyr, dtd); m3_mulv(t, Jrtd, u); m3_skew(sk, t);
        q_rot(u, corrected_delta_q, bcv);
        for (int k = 0; k < 3; k++) w[k] = fcv[k] + corrected_delta_p[k] - u[k];
        m3_mulv(b, sk, w);
        for (int k = 0; k < 3; k++) v[k] = dp_dsw[k] - a[k] + b[k];
        m3_mulv(t, Efw, v);
        col[0] = -t[0]; col[1] = -t[1]; col[2] = -t[2];
        /* -Jr_inv Exp(-raw_r) Exp(bcw) (Rcdq^-1 Jrtd lin_gyr dtd + Jr_drdsw dq_dsw) */
        double nr[3], E1[9], E2[9], Rcdq_inv[9], qci[4];
        v3_scale(nr, raw_r, -1.0); so3_exp_R(E1, nr); so3_exp_R(E2, bcw);
        q_inv(qci, corrected_delta_q); q_to_R(Rcdq_inv, qci);
        v3_scale(u, lin_gyr, dtd); m3_mulv(t, Jrtd, u); m3_mulv(a, Rcdq_inv, t);
        m3_mulv(b, Jr_drdsw, dq_dsw); v3_add(a, a, b);
        m3_mulv(t, E2, a); m3_mulv(u, E1, t); m3_mulv(t, Jr_dq_inv, u);
        col[3] = -t[0]; col[4] = -t[1]; col[5] = -t[2];
        mat_mul(jac[5], S, col, 6, 6, 1);
    }
    if (jac[6]) {
        double v[3], a[3], b[3], col[6];
        /* sv lin_vel - Rcdq sv vel_1 + skew(Jrtd sw lin_gyr) (fcv + cdp - Rcdq bcv) */
        m3_mulv(a, Rcdq, sv_v1);
        v3_scale(u, lin_gyr, sw); m3_mulv(t, Jrtd, u); m3_skew(sk, t);
        m3_mulv(u, Rcdq, bcv);
        for (int k = 0; k < 3; k++) w[k] = fcv[k] + corrected_delta_p[k] - u[k];
        m3_mulv(b, sk, w);
        for (int k = 0; k < 3; k++) v[k] = sv_lv[k] - a[k] + b[k];
        m3_mulv(t, Efw, v);
        col[0] = -t[0]; col[1] = -t[1]; col[2] = -t[2];
        /* -Jr_inv Exp(-raw_r) (Exp(bcw) Rcdq^-1 Jrtd sw lin_gyr - Jr_minus_td sw gyr_1) */
        double nr[3], E1[9], E2[9], Rcdq_inv[9], qci[4];
        v3_scale(nr, raw_r, -1.0); so3_exp_R(E1, nr); so3_exp_R(E2, bcw);
        q_inv(qci, corrected_delta_q); q_to_R(Rcdq_inv, qci);
        v3_scale(u, lin_gyr, sw); m3_mulv(t, Jrtd, u); m3_mulv(a, Rcdq_inv, t); m3_mulv(a, E2, a);
        v3_scale(u, gyr_1, sw); m3_mulv(b, Jr_minus_td, u);
        v3_sub(a, a, b); m3_mulv(u, E1, a); m3_mulv(t, Jr_dq_inv, u);
        col[3] = -t[0]; col[4] = -t[1]; col[5] = -t[2];
        mat_mul(jac[6], S, col, 6, 6, 1);
    }
    return 0;
}

/* ================================================================ plane factor */
/* PlaneFactor::Evaluate (factor/plane_factor.h:25-121) */
static int eval_plane(const viwb_globals *g, const double *c, const double *const *p, double *res, double **jac) {
    (void)c;
    const double *Pi = p[0], *Qi = p[0] + 3, *tio = p[1], *qio = p[1] + 3, *qpw = p[2];
    const double zpw = p[3][0];
    const double e3[3] = {0, 0, 1};
    double Ri[9], rio[9], Rpw[9], t[3], u[3], v[3], r[3];
    q_to_R(Ri, Qi); q_to_R(rio, qio); q_to_R(Rpw, qpw);
    m3_tmulv(t, Rpw, e3); m3_tmulv(u, Ri, t); m3_tmulv(v, rio, u);
    r[0] = v[0]; r[1] = v[1];
    q_rot(t, Qi, tio); v3_add(t, t, Pi); q_rot(u, qpw, t);
    r[2] = zpw + u[2];
    const double *w = g->plane_sqrt_info;
    res[0] = w[0] * r[0]; res[1] = w[1] * r[1]; res[2] = w[2] * r[2];
    if (!jac) return 0;
    double qi_inv[4], qpw_inv[4], qio_inv[4], sk[9], m[9];
    q_inv(qi_inv, Qi); q_inv(qpw_inv, qpw); q_inv(qio_inv, qio);
    if (jac[0]) {
        double *J = jac[0]; memset(J, 0, sizeof(double) * 21);
        double rioT[9];
        q_rot(t, qpw_inv, e3); q_rot(u, qi_inv, t); m3_skew(sk, u); m3_transpose(rioT, rio); m3_mul(m, rioT, sk);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) J[i * 7 + 3 + j] = m[i * 3 + j];
        for (int j = 0; j < 3; j++) J[2 * 7 + j] = Rpw[2 * 3 + j];                 /* e3^T Rpw */
        double RpwRi[9]; m3_mul(RpwRi, Rpw, Ri); m3_skew(sk, tio); m3_mul(m, RpwRi, sk);
        for (int j = 0; j < 3; j++) J[2 * 7 + 3 + j] = -m[2 * 3 + j];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 7; j++) J[i * 7 + j] *= w[i];
    }
    if (jac[1]) {
        double *J = jac[1]; memset(J, 0, sizeof(double) * 21);
        q_rot(t, qpw_inv, e3); q_rot(u, qi_inv, t); q_rot(v, qio_inv, u); m3_skew(sk, v);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) J[i * 7 + 3 + j] = sk[i * 3 + j];
        double RpwRi[9]; m3_mul(RpwRi, Rpw, Ri);
        for (int j = 0; j < 3; j++) J[2 * 7 + j] = RpwRi[2 * 3 + j];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 7; j++) J[i * 7 + j] *= w[i];
    }
    if (jac[2]) {
        double *J = jac[2]; memset(J, 0, sizeof(double) * 12);
        double rioT[9], RiT[9], a[9];
        q_rot(t, qpw_inv, e3); m3_skew(sk, t); m3_transpose(rioT, rio); m3_transpose(RiT, Ri);
        m3_mul(a, rioT, RiT); m3_mul(m, a, sk);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) J[i * 4 + j] = m[i * 3 + j];
        q_rot(t, Qi, tio); v3_add(t, t, Pi); m3_skew(sk, t); m3_mul(m, Rpw, sk);
        for (int j = 0; j < 3; j++) J[2 * 4 + j] = -m[2 * 3 + j];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) J[i * 4 + j] *= w[i];
    }
    if (jac[3]) { jac[3][0] = 0; jac[3][1] = 0; jac[3][2] = w[2] * 1.0; }
    return 0;
}

int vo_factor_evaluate(int factor_type, const viwb_globals *globals, const double *consts,
                       const double *const *parameters, double *residuals, double **jacobians) {
    switch (factor_type) {
    case VIWB_F_PROJ_2F1C: return eval_proj_2f1c(globals, consts, parameters, residuals, jacobians);
    case VIWB_F_PROJ_2F2C: return eval_proj_2f2c(globals, consts, parameters, residuals, jacobians);
    case VIWB_F_PROJ_1F2C: return eval_proj_1f2c(globals, consts, parameters, residuals, jacobians);
    case VIWB_F_IMU: return eval_imu(globals, consts, parameters, residuals, jacobians);
    case VIWB_F_WHEEL: return eval_wheel(globals, consts, parameters, residuals, jacobians);
    case VIWB_F_PLANE: return eval_plane(globals, consts, parameters, residuals, jacobians);
    }
    return VIWB_ERR_INVALID;
}

/* ================================================================ prior */
/* MarginalizationFactor::Evaluate (factor/marginalization_factor.cpp:349-397) */
int vo_prior_evaluate(const viwb_prior *prior, const double *state, double *residuals, double *jacobian) {
    int n = prior->n;
    double *dx = (double *)calloc(n, sizeof(double));
    for (int i = 0; i < prior->num_blocks; i++) {
        int b = prior->block_id[i], size = viwb_block_size(b), idx = prior->block_idx[i], off = viwb_block_offset(b);
        const double *x = state + off, *x0 = prior->x0 + off;
        if (size != 7) {
            for (int k = 0; k < size; k++) dx[idx + k] = x[k] - x0[k];
        } else {
            for (int k = 0; k < 3; k++) dx[idx + k] = x[k] - x0[k];
            double q0inv[4], dq[4];
            q_inv(q0inv, x0 + 3); q_mul(dq, q0inv, x + 3);
            for (int k = 0; k < 3; k++) dx[idx + 3 + k] = 2.0 * dq[k];
            if (!(dq[3] >= 0)) for (int k = 0; k < 3; k++) dx[idx + 3 + k] = 2.0 * -dq[k];
        }
    }
    for (int i = 0; i < n; i++) {
        double s = prior->r[i];
        for (int k = 0; k < n; k++) s += prior->J[i * n + k] * dx[k];
        residuals[i] = s;
    }
    free(dx);
    if (jacobian) {
        memset(jacobian, 0, sizeof(double) * n * VIWB_STATE_FIXED);
        for (int i = 0; i < prior->num_blocks; i++) {
            int b = prior->block_id[i], local = viwb_block_marg_size(b), idx = prior->block_idx[i], off = viwb_block_offset(b);
            for (int r = 0; r < n; r++) for (int k = 0; k < local; k++) jacobian[r * VIWB_STATE_FIXED + off + k] = prior->J[r * n + idx + k];
        }
    }
    return 0;
}

/* ceres::HuberLoss::Evaluate (ceres-solver loss_function.cc, third party) */
void vo_huber(double a, double s, double rho[3]) {
    double b = a * a;
    if (s > b) {
        double r = sqrt(s);
        rho[0] = 2.0 * a * r - b;
        rho[1] = a / r; if (rho[1] < 2.2250738585072014e-308) rho[1] = 2.2250738585072014e-308;
        rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

/* ================================================================ pre-integration */
/* IntegrationBase::midPointIntegration / propagate (factor/integration_base.h:63-167) */
void vo_imu_preintegrate(int n, const double *dt, const double *acc, const double *gyr, const double *ba,
                         const double *bg, const double *noise_sigma, double *rec) {
    double jacobian[225], covariance[225], noise[18], delta_p[3] = {0, 0, 0}, delta_q[4] = {0, 0, 0, 1}, delta_v[3] = {0, 0, 0};
    double sum_dt = 0;
    memset(jacobian, 0, sizeof jacobian); memset(covariance, 0, sizeof covariance);
    for (int i = 0; i < 15; i++) jacobian[i * 15 + i] = 1;
    double an = noise_sigma[0], gn = noise_sigma[1], aw = noise_sigma[2], gw = noise_sigma[3];
    for (int i = 0; i < 3; i++) { noise[i] = an * an; noise[3 + i] = gn * gn; noise[6 + i] = an * an; noise[9 + i] = gn * gn; noise[12 + i] = aw * aw; noise[15 + i] = gw * gw; }
    double *F = (double *)malloc(sizeof(double) * (225 + 270 + 225 + 225 + 270));
    double *V = F + 225, *T1 = V + 270, *T2 = T1 + 225, *VN = T2 + 225;
    for (int s = 0; s < n; s++) {
        const double _dt = dt[s], *a0 = acc + 3 * s, *g0 = gyr + 3 * s, *a1 = acc + 3 * (s + 1), *g1 = gyr + 3 * (s + 1);
        double t[3], un_acc_0[3], un_gyr[3], ddq[4], result_q[4], un_acc_1[3], un_acc[3], result_p[3], result_v[3];
        v3_sub(t, a0, ba); q_rot(un_acc_0, delta_q, t);
        for (int k = 0; k < 3; k++) un_gyr[k] = 0.5 * (g0[k] + g1[k]) - bg[k];
        ddq[3] = 1; ddq[0] = un_gyr[0] * _dt / 2; ddq[1] = un_gyr[1] * _dt / 2; ddq[2] = un_gyr[2] * _dt / 2;
        q_mul(result_q, delta_q, ddq);
        v3_sub(t, a1, ba); q_rot(un_acc_1, result_q, t);
        for (int k = 0; k < 3; k++) {
            un_acc[k] = 0.5 * (un_acc_0[k] + un_acc_1[k]);
            result_p[k] = delta_p[k] + delta_v[k] * _dt + 0.5 * un_acc[k] * _dt * _dt;
            result_v[k] = delta_v[k] + un_acc[k] * _dt;
        }
        /* update_jacobian */
        double w_x[3], a_0_x[3], a_1_x[3], R_w_x[9], R_a_0_x[9], R_a_1_x[9], Rd[9], Rr[9], I[9], IwR[9], m[9], m2[9];
        for (int k = 0; k < 3; k++) w_x[k] = 0.5 * (g0[k] + g1[k]) - bg[k];
        v3_sub(a_0_x, a0, ba); v3_sub(a_1_x, a1, ba);
        m3_skew(R_w_x, w_x); m3_skew(R_a_0_x, a_0_x); m3_skew(R_a_1_x, a_1_x);
        q_to_R(Rd, delta_q); q_to_R(Rr, result_q); m3_identity(I);
        m3_scale(m, R_w_x, _dt); m3_sub(IwR, I, m);                        /* I - R_w_x dt */
        memset(F, 0, sizeof(double) * 225); memset(V, 0, sizeof(double) * 270);
        double Rr_Ra1[9], Rr_Ra1_IwR[9], Rd_Ra0[9];
        m3_mul(Rd_Ra0, Rd, R_a_0_x); m3_mul(Rr_Ra1, Rr, R_a_1_x); m3_mul(Rr_Ra1_IwR, Rr_Ra1, IwR);
        put33(F, 15, 0, 0, I);
        for (int k = 0; k < 9; k++) m[k] = -0.25 * Rd_Ra0[k] * _dt * _dt + -0.25 * Rr_Ra1_IwR[k] * _dt * _dt;
        put33(F, 15, 0, 3, m);
        m3_scale(m, I, _dt); put33(F, 15, 0, 6, m);
        for (int k = 0; k < 9; k++) m[k] = -0.25 * (Rd[k] + Rr[k]) * _dt * _dt;
        put33(F, 15, 0, 9, m);
        for (int k = 0; k < 9; k++) m[k] = -0.25 * Rr_Ra1[k] * _dt * _dt * -_dt;
        put33(F, 15, 0, 12, m);
        put33(F, 15, 3, 3, IwR);
        m3_scale(m, I, -1.0 * _dt); put33(F, 15, 3, 12, m);
        for (int k = 0; k < 9; k++) m[k] = -0.5 * Rd_Ra0[k] * _dt + -0.5 * Rr_Ra1_IwR[k] * _dt;
        put33(F, 15, 6, 3, m);
        put33(F, 15, 6, 6, I);
        for (int k = 0; k < 9; k++) m[k] = -0.5 * (Rd[k] + Rr[k]) * _dt;
        put33(F, 15, 6, 9, m);
        for (int k = 0; k < 9; k++) m[k] = -0.5 * Rr_Ra1[k] * _dt * -_dt;
        put33(F, 15, 6, 12, m);
        put33(F, 15, 9, 9, I); put33(F, 15, 12, 12, I);
        m3_scale(m, Rd, 0.25 * _dt * _dt); put33(V, 18, 0, 0, m);
        for (int k = 0; k < 9; k++) m2[k] = 0.25 * -Rr_Ra1[k] * _dt * _dt * 0.5 * _dt;
        put33(V, 18, 0, 3, m2); put33(V, 18, 0, 9, m2);
        m3_scale(m, Rr, 0.25 * _dt * _dt); put33(V, 18, 0, 6, m);
        m3_scale(m, I, 0.5 * _dt); put33(V, 18, 3, 3, m); put33(V, 18, 3, 9, m);
        m3_scale(m, Rd, 0.5 * _dt); put33(V, 18, 6, 0, m);
        for (int k = 0; k < 9; k++) m2[k] = 0.5 * -Rr_Ra1[k] * _dt * 0.5 * _dt;
        put33(V, 18, 6, 3, m2); put33(V, 18, 6, 9, m2);
        m3_scale(m, Rr, 0.5 * _dt); put33(V, 18, 6, 6, m);
        m3_scale(m, I, _dt); put33(V, 18, 9, 12, m); put33(V, 18, 12, 15, m);
        /* jacobian = F * jacobian; covariance = F cov F^T + V noise V^T */
        mat_mul(T1, F, jacobian, 15, 15, 15); memcpy(jacobian, T1, sizeof(double) * 225);
        mat_mul(T1, F, covariance, 15, 15, 15);
        for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) {
            double sF = 0; for (int k = 0; k < 15; k++) sF += T1[i * 15 + k] * F[j * 15 + k];
            T2[i * 15 + j] = sF;
        }
        for (int i = 0; i < 15; i++) for (int k = 0; k < 18; k++) VN[i * 18 + k] = V[i * 18 + k] * noise[k];
        for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) {
            double sV = 0; for (int k = 0; k < 18; k++) sV += VN[i * 18 + k] * V[j * 18 + k];
            covariance[i * 15 + j] = T2[i * 15 + j] + sV;
        }
        v3_copy(delta_p, result_p); memcpy(delta_q, result_q, sizeof result_q); v3_copy(delta_v, result_v);
        q_normalize(delta_q);
        sum_dt += _dt;
    }
    free(F);
    rec[0] = sum_dt; v3_copy(rec + 1, delta_p); memcpy(rec + 4, delta_q, 4 * sizeof(double)); v3_copy(rec + 8, delta_v);
    v3_copy(rec + 11, ba); v3_copy(rec + 14, bg);
    const int blk[5][2] = {{0, 9}, {0, 12}, {3, 12}, {6, 9}, {6, 12}};   /* dp_dba dp_dbg dq_dbg dv_dba dv_dbg */
    for (int b = 0; b < 5; b++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        rec[17 + 9 * b + i * 3 + j] = jacobian[(blk[b][0] + i) * 15 + blk[b][1] + j];
    memcpy(rec + 62, covariance, sizeof covariance);
}

/* WheelIntegrationBase::midPointIntegration / propagate (factor/wheel_integration_base.h:67-177) */
void vo_wheel_preintegrate(int n, const double *dt, const double *vel, const double *gyr, const double *s,
                           double td, const double *noise_sigma, double *rec) {
    double jac[18], cov[36], noise[12], delta_p[3] = {0, 0, 0}, delta_q[4] = {0, 0, 0, 1}, sum_dt = 0;
    memset(jac, 0, sizeof jac); memset(cov, 0, sizeof cov);
    double vn = noise_sigma[0], gn = noise_sigma[1];
    for (int i = 0; i < 3; i++) { noise[i] = vn * vn; noise[3 + i] = gn * gn; noise[6 + i] = vn * vn; noise[9 + i] = gn * gn; }
    const double sx = s[0], sy = s[1], sw = s[2];
    const double sv[3] = {sx, sy, 1.0};
    for (int st = 0; st < n; st++) {
        const double _dt = dt[st], *v0 = vel + 3 * st, *g0 = gyr + 3 * st, *v1 = vel + 3 * (st + 1), *g1 = gyr + 3 * (st + 1);
        double sv_v0[3] = {sv[0] * v0[0], sv[1] * v0[1], sv[2] * v0[2]}, sv_v1[3] = {sv[0] * v1[0], sv[1] * v1[1], sv[2] * v1[2]};
        double un_vel_0[3], un_gyr[3], ddq[4], result_q[4], un_vel_1[3], result_p[3];
        q_rot(un_vel_0, delta_q, sv_v0);
        for (int k = 0; k < 3; k++) un_gyr[k] = 0.5 * sw * (g0[k] + g1[k]);
        ddq[3] = 1; ddq[0] = un_gyr[0] * _dt / 2; ddq[1] = un_gyr[1] * _dt / 2; ddq[2] = un_gyr[2] * _dt / 2;
        q_mul(result_q, delta_q, ddq);
        q_rot(un_vel_1, result_q, sv_v1);
        for (int k = 0; k < 3; k++) result_p[k] = delta_p[k] + 0.5 * (un_vel_0[k] + un_vel_1[k]) * _dt;
        double R_v0[9], R_v1[9], Rd[9], Rr[9], Rdd[9], RddT[9], Jr[9], F[36], V[72], m[9], m2[9], t3[3];
        m3_skew(R_v0, sv_v0); m3_skew(R_v1, sv_v1);
        q_to_R(Rd, delta_q); q_to_R(Rr, result_q); q_to_R(Rdd, ddq); m3_transpose(RddT, Rdd);
        memset(F, 0, sizeof F); memset(V, 0, sizeof V);
        m3_identity(m); put33(F, 6, 0, 0, m);
        m3_mul(m, Rd, R_v0); m3_mul(m2, Rr, R_v1); m3_mul(m2, m2, RddT);
        for (int k = 0; k < 9; k++) m[k] = -0.5 * _dt * (m[k] + m2[k]);
        put33(F, 6, 0, 3, m); put33(F, 6, 3, 3, RddT);
        v3_scale(t3, un_gyr, _dt); so3_Jr(Jr, t3);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i * 3 + j] = 0.5 * _dt * Rd[i * 3 + j] * sv[j];
        put33(V, 12, 0, 0, m);
        m3_mul(m2, Rr, R_v1); m3_mul(m2, m2, Jr); m3_scale(m2, m2, -0.25 * _dt * _dt);
        put33(V, 12, 0, 3, m2); put33(V, 12, 0, 9, m2);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i * 3 + j] = 0.5 * _dt * Rr[i * 3 + j] * sv[j];
        put33(V, 12, 0, 6, m);
        m3_scale(m, Jr, 0.5 * sw * _dt); put33(V, 12, 3, 3, m); put33(V, 12, 3, 9, m);
        /* jacobian w.r.t. intrinsics */
        for (int k = 0; k < 3; k++) {
            jac[k * 3 + 0] += 0.5 * (Rd[k * 3 + 0] * v0[0] + Rr[k * 3 + 0] * v1[0]) * _dt;
            jac[k * 3 + 1] += 0.5 * (Rd[k * 3 + 1] * v0[1] + Rr[k * 3 + 1] * v1[1]) * _dt;
        }
        double dr_last[3] = {jac[9 + 2], jac[12 + 2], jac[15 + 2]}, gm[3], Jg[3], dr_new[3], a[3], b[3], sk[9];
        for (int k = 0; k < 3; k++) gm[k] = 0.5 * (g0[k] + g1[k]) * _dt;
        m3_mulv(Jg, Jr, gm);
        for (int k = 0; k < 3; k++) { jac[(3 + k) * 3 + 2] += Jg[k]; dr_new[k] = jac[(3 + k) * 3 + 2]; }
        m3_skew(sk, dr_last); m3_mulv(a, sk, sv_v0); m3_mulv(a, Rd, a);
        m3_skew(sk, dr_new); m3_mulv(b, sk, sv_v1); m3_mulv(b, Rr, b);
        for (int k = 0; k < 3; k++) jac[k * 3 + 2] += 0.5 * (a[k] + b[k]) * _dt;
        double T1[36], T2[36], VN[72];
        mat_mul(T1, F, cov, 6, 6, 6);
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double sF = 0; for (int k = 0; k < 6; k++) sF += T1[i * 6 + k] * F[j * 6 + k]; T2[i * 6 + j] = sF; }
        for (int i = 0; i < 6; i++) for (int k = 0; k < 12; k++) VN[i * 12 + k] = V[i * 12 + k] * noise[k];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double sV = 0; for (int k = 0; k < 12; k++) sV += VN[i * 12 + k] * V[j * 12 + k]; cov[i * 6 + j] = T2[i * 6 + j] + sV; }
        v3_copy(delta_p, result_p); memcpy(delta_q, result_q, sizeof result_q); q_normalize(delta_q);
        sum_dt += _dt;
    }
    v3_copy(rec, delta_p); memcpy(rec + 3, delta_q, 4 * sizeof(double));
    memcpy(rec + 7, jac, sizeof jac); memcpy(rec + 25, cov, sizeof cov);
    rec[61] = sx; rec[62] = sy; rec[63] = sw; rec[64] = td;
    v3_copy(rec + 65, vel); v3_copy(rec + 68, gyr);
    v3_copy(rec + 71, vel + 3 * n); v3_copy(rec + 74, gyr + 3 * n);
    rec[77] = sum_dt;
}

void vo_default_options(viwb_options *o) {
    o->max_num_iterations = 8; o->max_solver_time_in_seconds = 0;
    o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
    o->max_num_consecutive_invalid_steps = 5; o->jacobi_scaling = 1;
}
void vo_default_globals(viwb_globals *g) {
    g->G[0] = 0; g->G[1] = 0; g->G[2] = 9.81007;                    /* config/euroc/euroc_mono_imu_config.yaml g_norm */
    g->vis_sqrt_info[0] = 460.0 / 1.5; g->vis_sqrt_info[1] = 0; g->vis_sqrt_info[2] = 0; g->vis_sqrt_info[3] = 460.0 / 1.5;
    g->plane_sqrt_info[0] = 1.0 / 0.01; g->plane_sqrt_info[1] = 1.0 / 0.01; g->plane_sqrt_info[2] = 1.0 / 0.05;
    g->huber_delta = 1.0;
}

// factors.cuh -- per-factor residual + tangent-space Jacobian evaluation (FP64), one thread per factor.
//
// What is computed is fixed by the reference (paths relative to /root/reference/vins_estimator/src):
//   vis_eval   : ProjectionTwoFrameOneCamFactor / TwoFrameTwoCam / OneFrameTwoCam ::Evaluate
//                (factor/projectionTwoFrameOneCamFactor.cpp:45-152, projectionTwoFrameTwoCamFactor.cpp:43-166,
//                 projectionOneFrameTwoCamFactor.cpp:42-134), one fused routine for the three types
//   imu_eval   : IMUFactor::Evaluate (factor/imu_factor.h:30-192) + IntegrationBase::evaluate (integration_base.h:169-195)
//   wheel_eval : WheelFactor::Evaluate (factor/wheel_factor.h:28-246) + WheelIntegrationBase::evaluate (:179-218)
//   plane_eval : PlaneFactor::Evaluate (factor/plane_factor.h:25-121)
// How it is computed is ours: rotation-matrix chains shared between residual and Jacobian, Jacobians emitted
// directly in the 6-dof tangent layout the normal equations consume (the zero 7th column of the reference's
// 2x7 blocks is never materialised), whitening matrices precomputed once per solve instead of per Evaluate.
#pragma once
#include "vmath.cuh"

namespace viwb {

// ---------------------------------------------------------------------------------------------- visual
struct VisOut {
    double r[2];
    double JA[12];    // 2x6 d r / d pose_i   (TwoFrame*)          [dp(3) dtheta(3)]
    double JB[12];    // 2x6 d r / d pose_j   (TwoFrame*)
    double JE0[12];   // 2x6 d r / d ex_pose0
    double JE1[12];   // 2x6 d r / d ex_pose1 (TwoCam)
    double Jl[2];     // d r / d inverse depth
    double Jtd[2];    // d r / d td
};

// 2x3 * [L | R] (3x3 each) -> 2x6 row-major
VIWB_HD void put26(double *J, const double *red, const M3 &L, const M3 &R) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            J[r * 6 + c] = red[r * 3] * L.m[c] + red[r * 3 + 1] * L.m[3 + c] + red[r * 3 + 2] * L.m[6 + c];
            J[r * 6 + 3 + c] = red[r * 3] * R.m[c] + red[r * 3 + 1] * R.m[3 + c] + red[r * 3 + 2] * R.m[6 + c];
        }
}
VIWB_HD void red_mulv(double *o, const double *red, const V3 &v, double s) {
    o[0] = (red[0] * v.x + red[1] * v.y + red[2] * v.z) * s;
    o[1] = (red[3] * v.x + red[4] * v.y + red[5] * v.z) * s;
}

// type: 0 = 2F1C, 1 = 2F2C, 2 = 1F2C.  obs: pts_i(3) pts_j(3) vel_i(2) vel_j(2) td_i td_j.
// S = sqrt_info (2x2 row-major).  pose pointers are [p(3), q(x,y,z,w)].
// COMMON = false leaves JE0 / JE1 / Jtd unwritten (windows whose extrinsics and td are constant never read them)
template <bool COMMON>
VIWB_HD void vis_eval_t(int type, const double *obs, const double *pose_i, const double *pose_j, const double *ex0,
                        const double *ex1, double inv_dep, double td, const double *S, bool want_j, VisOut &o) {
    const V3 pts_i = ld3(obs), pts_j = ld3(obs + 3);
    const V3 vel_i = v3(obs[6], obs[7], 0.0), vel_j = v3(obs[8], obs[9], 0.0);
    const V3 pi_td = pts_i - (td - obs[10]) * vel_i;
    const V3 pj_td = pts_j - (td - obs[11]) * vel_j;
    const double inv_l = 1.0 / inv_dep;
    const M3 Ric = qR(ldq(ex0 + 3));
    const V3 tic = ld3(ex0);
    const V3 Xci = pi_td * inv_l;
    const V3 Xbi = Ric * Xci + tic;
    M3 Ri, Rj, RcT;          // RcT = (rotation of the camera the point is projected into)^T
    V3 Xbj, tcj;
    if (type == 2) {
        Xbj = Xbi;
        RcT = transpose(qR(ldq(ex1 + 3))); tcj = ld3(ex1);
        Ri = m3_identity(); Rj = Ri;
    } else {
        Ri = qR(ldq(pose_i + 3)); Rj = qR(ldq(pose_j + 3));
        const V3 Xw = Ri * Xbi + ld3(pose_i);
        Xbj = tmul(Rj, Xw - ld3(pose_j));
        if (type == 1) { RcT = transpose(qR(ldq(ex1 + 3))); tcj = ld3(ex1); }
        else { RcT = transpose(Ric); tcj = tic; }
    }
    const V3 Xcj = RcT * (Xbj - tcj);
    const double iz = 1.0 / Xcj.z;
    const double e0 = Xcj.x * iz - pj_td.x, e1 = Xcj.y * iz - pj_td.y;
    o.r[0] = S[0] * e0 + S[1] * e1;
    o.r[1] = S[2] * e0 + S[3] * e1;
    if (!want_j) return;
    // reduce = S * [1/z 0 -x/z^2 ; 0 1/z -y/z^2]
    double red[6];
    {
        const double a = iz, b = -Xcj.x * iz * iz, c = -Xcj.y * iz * iz;
        red[0] = S[0] * a; red[1] = S[1] * a; red[2] = S[0] * b + S[1] * c;
        red[3] = S[2] * a; red[4] = S[3] * a; red[5] = S[2] * b + S[3] * c;
    }
    const double sv0 = S[0] * vel_j.x + S[1] * vel_j.y, sv1 = S[2] * vel_j.x + S[3] * vel_j.y;
    if (type == 2) {
        const M3 B = RcT * Ric;                                   // ric2^T ric
        if (COMMON) {
            put26(o.JE0, red, RcT, -(B * skew(Xci)));
            put26(o.JE1, red, -RcT, skew(Xcj));
            red_mulv(o.Jtd, red, B * vel_i, -inv_l);
            o.Jtd[0] += sv0; o.Jtd[1] += sv1;
        }
        red_mulv(o.Jl, red, B * pts_i, -inv_l * inv_l);           // quirk 3: un-compensated pts_i (:119)
        for (int k = 0; k < 12; k++) { o.JA[k] = 0.0; o.JB[k] = 0.0; }
        return;
    }
    const M3 M = RcT * transpose(Rj);                             // rc^T Rj^T
    const M3 MRi = M * Ri;
    const M3 T = MRi * Ric;
    put26(o.JA, red, M, -(MRi * skew(Xbi)));
    put26(o.JB, red, -M, RcT * skew(Xbj));
    if (COMMON) {
        if (type == 0) {
            const M3 L = RcT * (transpose(Rj) * Ri - m3_identity());
            const V3 w = RcT * (tmul(Rj, Ri * tic + ld3(pose_i) - ld3(pose_j)) - tic);
            put26(o.JE0, red, L, -(T * skew(Xci)) + skew(T * Xci) + skew(w));
            for (int k = 0; k < 12; k++) o.JE1[k] = 0.0;
        } else {
            put26(o.JE0, red, MRi, -(T * skew(Xci)));
            put26(o.JE1, red, -RcT, skew(Xcj));
        }
        red_mulv(o.Jtd, red, T * vel_i, -inv_l);
        o.Jtd[0] += sv0; o.Jtd[1] += sv1;
    }
    red_mulv(o.Jl, red, T * pi_td, -inv_l * inv_l);
}
VIWB_HD void vis_eval(int type, const double *obs, const double *pose_i, const double *pose_j, const double *ex0,
                      const double *ex1, double inv_dep, double td, const double *S, bool want_j, VisOut &o) {
    vis_eval_t<true>(type, obs, pose_i, pose_j, ex0, ex1, inv_dep, td, S, want_j, o);
}

// ceres::HuberLoss + Corrector for a 2-row residual block: rho'' < 0 whenever s > delta^2, so the corrector always
// takes the simple sqrt(rho') scaling (marginalization_factor.cpp:46-57).  Returns the scale and writes cost = rho/2.
VIWB_HD double huber_scale(double s, double delta, double &half_rho) {
    const double b = delta * delta;
    if (s > b) { const double r = sqrt(s); half_rho = 0.5 * (2.0 * delta * r - b); return sqrt(delta / r); }
    half_rho = 0.5 * s;
    return 1.0;
}

// ---------------------------------------------------------------------------------------------- IMU
// Upper-triangular S with S^T S = cov^-1  (== LLT(cov.inverse()).matrixL().transpose(), imu_factor.h:75):
// reverse Cholesky cov = U U^T (U upper), S = U^-1.  n <= 15.  Returns false if cov is not positive definite.
VIWB_HD bool sqrt_info_upper(int n, const double *cov, double *S) {
    double U[225];
    for (int j = n - 1; j >= 0; j--) {
        double d = cov[j * n + j];
        for (int k = j + 1; k < n; k++) d -= U[j * n + k] * U[j * n + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d); U[j * n + j] = d;
        for (int i = 0; i < j; i++) {
            double s = cov[i * n + j];
            for (int k = j + 1; k < n; k++) s -= U[i * n + k] * U[j * n + k];
            U[i * n + j] = s / d;
        }
    }
    // S = U^-1 (upper): column by column back substitution
    for (int i = 0; i < n * n; i++) S[i] = 0.0;
    for (int c = 0; c < n; c++) {
        S[c * n + c] = 1.0 / U[c * n + c];
        for (int i = c - 1; i >= 0; i--) {
            double s = 0.0;
            for (int k = i + 1; k <= c; k++) s += U[i * n + k] * S[k * n + c];
            S[i * n + c] = -s / U[i * n + i];
        }
    }
    return true;
}

VIWB_HD void put33(double *J, int ld, int r0, int c0, const M3 &m) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[(r0 + i) * ld + c0 + j] = m.m[i * 3 + j];
}
// J (rows x cols, row-major) <- S J with S upper triangular rows x rows
VIWB_HD void whiten_upper(int rows, int cols, const double *S, double *J) {
    for (int c = 0; c < cols; c++)
        for (int i = 0; i < rows; i++) {
            double s = 0.0;
            for (int k = i; k < rows; k++) s += S[i * rows + k] * J[k * cols + c];
            J[i * cols + c] = s;     // row i only needs rows >= i, which are still un-whitened
        }
}

// rec = the 287-double record of include/viwb.h; S = 15x15 upper sqrt-info (nullptr: leave r and J un-whitened).  Outputs the whitened residual r[15]
// and, if want_j, whitened tangent Jacobians J (15 x 30 row-major: pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9).
VIWB_HD void imu_eval(const double *rec, const double *S, const double *G, const double *pose_i, const double *sb_i,
                      const double *pose_j, const double *sb_j, bool want_j, double *r, double *J) {
    const double dt = rec[0];
    const V3 dp = ld3(rec + 1), dv = ld3(rec + 8), lin_ba = ld3(rec + 11), lin_bg = ld3(rec + 14);
    const Q4 dq = ldq(rec + 4);
    const M3 dp_dba = m3_ld(rec + 17), dp_dbg = m3_ld(rec + 26), dq_dbg = m3_ld(rec + 35), dv_dba = m3_ld(rec + 44), dv_dbg = m3_ld(rec + 53);
    const V3 Pi = ld3(pose_i), Pj = ld3(pose_j), Vi = ld3(sb_i), Vj = ld3(sb_j);
    const V3 Bai = ld3(sb_i + 3), Bgi = ld3(sb_i + 6), Baj = ld3(sb_j + 3), Bgj = ld3(sb_j + 6);
    const Q4 Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3);
    const V3 g = ld3(G);
    const V3 dba = Bai - lin_ba, dbg = Bgi - lin_bg;
    const Q4 cdq = dq * q_delta(dq_dbg * dbg);
    const V3 cdv = dv + dv_dba * dba + dv_dbg * dbg;
    const V3 cdp = dp + dp_dba * dba + dp_dbg * dbg;
    const Q4 Qi_inv = qinv(Qi);
    const M3 RiT = qR(Qi_inv);
    const V3 a_p = RiT * (0.5 * dt * dt * g + Pj - Pi - dt * Vi);
    const V3 a_v = RiT * (dt * g + Vj - Vi);
    const Q4 qij = Qi_inv * Qj;
    const Q4 qe = qinv(cdq) * qij;
    double raw[15];
    st3(raw, a_p - cdp);
    raw[3] = 2.0 * qe.x; raw[4] = 2.0 * qe.y; raw[5] = 2.0 * qe.z;
    st3(raw + 6, a_v - cdv); st3(raw + 9, Baj - Bai); st3(raw + 12, Bgj - Bgi);
    if (S) { for (int i = 0; i < 15; i++) { double s = 0.0; for (int k = i; k < 15; k++) s += S[i * 15 + k] * raw[k]; r[i] = s; } }
    else for (int i = 0; i < 15; i++) r[i] = raw[i];
    if (!want_j) return;
    for (int i = 0; i < 450; i++) J[i] = 0.0;
    // pose_i (cols 0..5)
    put33(J, 30, 0, 0, -RiT);
    put33(J, 30, 0, 3, skew(a_p));
    {   // -(Qleft(Qj^-1 Qi) Qright(cdq)).bottomRightCorner<3,3>()
        const Q4 qji = qinv(Qj) * Qi;
        const M3 c = q_left33(qji) * q_right33(cdq) - outer(qvec(qji), qvec(cdq));
        put33(J, 30, 3, 3, -c);
    }
    put33(J, 30, 6, 3, skew(a_v));
    // speed-bias_i (cols 6..14)
    put33(J, 30, 0, 6, RiT * (-dt));
    put33(J, 30, 0, 9, -dp_dba);
    put33(J, 30, 0, 12, -dp_dbg);
    put33(J, 30, 3, 12, -(q_left33(qinv(Qj) * Qi * dq) * dq_dbg));     // quirk 5: un-corrected delta_q (imu_factor.h:138)
    put33(J, 30, 6, 6, -RiT);
    put33(J, 30, 6, 9, -dv_dba);
    put33(J, 30, 6, 12, -dv_dbg);
    put33(J, 30, 9, 9, -m3_identity());
    put33(J, 30, 12, 12, -m3_identity());
    // pose_j (cols 15..20)
    put33(J, 30, 0, 15, RiT);
    put33(J, 30, 3, 18, q_left33(qe));                                   // Qleft(cdq^-1 Qi^-1 Qj)
    // speed-bias_j (cols 21..29)
    put33(J, 30, 6, 21, RiT);
    put33(J, 30, 9, 24, m3_identity());
    put33(J, 30, 12, 27, m3_identity());
    if (S) whiten_upper(15, 30, S, J);
}

// ---------------------------------------------------------------------------------------------- wheel
// rec = the 78-double record; S = 6x6 upper sqrt-info.  J (6 x 22): pose_i 6 | pose_j 6 | ex_wheel 6 | sx | sy | sw | td_wheel.
VIWB_HD void wheel_eval(const double *rec, const double *S, const double *pose_i, const double *pose_j, const double *exw,
                        double sx, double sy, double sw, double td, bool want_j, double *r, double *J) {
    const V3 delta_p = ld3(rec); const Q4 delta_q = ldq(rec + 3);
    const double *Jp = rec + 7;
    const V3 dp_dsx = v3(Jp[0], Jp[3], Jp[6]), dp_dsy = v3(Jp[1], Jp[4], Jp[7]), dp_dsw = v3(Jp[2], Jp[5], Jp[8]);
    const V3 dq_dsw = v3(Jp[11], Jp[14], Jp[17]);
    const double lin_sx = rec[61], lin_sy = rec[62], lin_sw = rec[63], lin_td = rec[64];
    const V3 lin_vel = ld3(rec + 65), lin_gyr = ld3(rec + 68), vel_1 = ld3(rec + 71), gyr_1 = ld3(rec + 74);
    const V3 Pi = ld3(pose_i), Pj = ld3(pose_j), tio = ld3(exw);
    const Q4 Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3), qio = ldq(exw + 3);
    const M3 Ri = qR(Qi), Rj = qR(Qj), rio = qR(qio);
    const double dsx = sx - lin_sx, dsy = sy - lin_sy, dsw = sw - lin_sw, dtd = td - lin_td;
    const V3 cdp = delta_p + dp_dsx * dsx + dp_dsy * dsy + dp_dsw * dsw;
    const Q4 cdq = qnormalized(delta_q) * so3_exp_q(dq_dsw * dsw);
    const V3 fcw = lin_gyr * (sw * dtd);
    const V3 bcw = gyr_1 * (sw * dtd);
    const Q4 qfw = so3_exp_q(fcw);
    const Q4 dq_time = qfw * qnormalized(cdq) * so3_exp_q(-bcw);
    const V3 sv_lv = v3(sx * lin_vel.x, sy * lin_vel.y, lin_vel.z), sv_v1 = v3(sx * vel_1.x, sy * vel_1.y, vel_1.z);
    const V3 fcv = sv_lv * dtd, bcv = sv_v1 * dtd;
    const M3 Efw = qR(qfw);
    const V3 dp_time = Efw * (fcv + cdp - qrot(cdq, bcv));
    const M3 Rwo = Ri * rio;
    const V3 d = Rj * tio + Pj - Ri * tio - Pi;
    const Q4 q_iio = Qi * qio, q_iio_inv = qinv(q_iio);
    double raw[6];
    st3(raw, tmul(Rwo, d) - dp_time);
    const V3 rth = so3_log_q(qinv(dq_time) * q_iio_inv * Qj * qio);
    st3(raw + 3, rth);
    if (S) { for (int i = 0; i < 6; i++) { double s = 0.0; for (int k = i; k < 6; k++) s += S[i * 6 + k] * raw[k]; r[i] = s; } }
    else for (int i = 0; i < 6; i++) r[i] = raw[i];
    if (!want_j) return;
    for (int i = 0; i < 132; i++) J[i] = 0.0;
    const M3 Jri = so3_Jr_inv(rth);
    const M3 Jr_drdsw = so3_Jr(dq_dsw * (sw - lin_sw));
    const M3 R_iio_inv = qR(q_iio_inv);
    // pose_i
    put33(J, 22, 0, 0, -R_iio_inv);
    put33(J, 22, 0, 3, transpose(Rwo) * (Ri * skew(tio)) + transpose(rio) * skew(tmul(Ri, d)));
    put33(J, 22, 3, 3, -(Jri * qR(qinv(Qj * qio) * Qi)));
    // pose_j
    put33(J, 22, 0, 6, R_iio_inv);
    put33(J, 22, 0, 9, -(qR(q_iio_inv * Qj) * skew(tio)));
    put33(J, 22, 3, 9, Jri * qR(qinv(qio)));
    // ex_wheel
    put33(J, 22, 0, 12, R_iio_inv * (Rj - Ri));
    put33(J, 22, 0, 15, skew(qrot(q_iio_inv, qrot(Qj, tio) + Pj - qrot(Qi, tio) - Pi)));
    put33(J, 22, 3, 15, Jri * (m3_identity() - qR(qinv(Qj * qio) * Qi * qio)));
    const M3 Jrtd = so3_Jr(fcw), Jr_minus_td = so3_Jr(-fcw);
    const M3 Rcdq = qR(cdq), Rcdq_inv = qR(qinv(cdq));
    const M3 Efv = so3_exp_R(fcv);                                   // quirk 4: Exp of a velocity*dt vector (wheel_factor.h:198,210)
    {   // sx, sy (position rows only)
        const V3 tx = Efv * (v3(lin_vel.x * dtd, 0, 0) + dp_dsx - Rcdq * v3(vel_1.x * dtd, 0, 0));
        const V3 ty = Efv * (v3(0, lin_vel.y * dtd, 0) + dp_dsy - Rcdq * v3(0, vel_1.y * dtd, 0));
        for (int k = 0; k < 3; k++) { J[k * 22 + 18] = -comp(tx, k); J[k * 22 + 19] = -comp(ty, k); }
    }
    const M3 E1 = so3_exp_R(-rth), E2 = so3_exp_R(bcw);
    {   // sw
        const V3 a = Rcdq * (skew(Jr_drdsw * dq_dsw) * (sv_v1 * dtd));
        const V3 b = skew(Jrtd * (lin_gyr * dtd)) * (fcv + cdp - qrot(cdq, bcv));
        const V3 tp = Efw * (dp_dsw - a + b);
        const V3 tr = Jri * (E1 * (E2 * (Rcdq_inv * (Jrtd * (lin_gyr * dtd)) + Jr_drdsw * dq_dsw)));
        for (int k = 0; k < 3; k++) { J[k * 22 + 20] = -comp(tp, k); J[(3 + k) * 22 + 20] = -comp(tr, k); }
    }
    {   // td_wheel
        const V3 b = skew(Jrtd * (lin_gyr * sw)) * (fcv + cdp - Rcdq * bcv);
        const V3 tp = Efw * (sv_lv - Rcdq * sv_v1 + b);
        const V3 tr = Jri * (E1 * (E2 * (Rcdq_inv * (Jrtd * (lin_gyr * sw))) - Jr_minus_td * (gyr_1 * sw)));
        for (int k = 0; k < 3; k++) { J[k * 22 + 21] = -comp(tp, k); J[(3 + k) * 22 + 21] = -comp(tr, k); }
    }
    if (S) whiten_upper(6, 22, S, J);
}

// ---------------------------------------------------------------------------------------------- plane
// J (3 x 16): pose_i 6 | ex_wheel 6 | plane_R 3 | plane_Z 1.  w = diag sqrt-info (pitch, roll, zpw).
VIWB_HD void plane_eval(const double *w, const double *pose_i, const double *exw, const double *qpw_p, double zpw,
                        bool want_j, double *r, double *J) {
    const V3 Pi = ld3(pose_i), tio = ld3(exw), e3 = v3(0, 0, 1);
    const M3 Ri = qR(ldq(pose_i + 3)), rio = qR(ldq(exw + 3)), Rpw = qR(ldq(qpw_p));
    const V3 n_w = tmul(Rpw, e3);            // Rpw^T e3
    const V3 n_b = tmul(Ri, n_w);            // Ri^T Rpw^T e3
    const V3 n_o = tmul(rio, n_b);
    const V3 pw = Pi + Ri * tio;
    const V3 rp = Rpw * pw;
    r[0] = w[0] * n_o.x; r[1] = w[1] * n_o.y; r[2] = w[2] * (zpw + rp.z);
    if (!want_j) return;
    for (int i = 0; i < 48; i++) J[i] = 0.0;
    const M3 A = transpose(rio) * skew(n_b);                     // rows 0-1 of d/d theta_i
    const M3 RpwRi = Rpw * Ri;
    const M3 B = RpwRi * skew(tio);
    const M3 Cx = skew(n_o);
    const M3 Dq = transpose(rio) * (transpose(Ri) * skew(n_w));
    const M3 Eq = Rpw * skew(pw);
    for (int j = 0; j < 3; j++) {
        J[0 * 16 + 3 + j] = w[0] * A.m[j]; J[1 * 16 + 3 + j] = w[1] * A.m[3 + j];
        J[2 * 16 + j] = w[2] * Rpw.m[6 + j]; J[2 * 16 + 3 + j] = -w[2] * B.m[6 + j];
        J[0 * 16 + 9 + j] = w[0] * Cx.m[j]; J[1 * 16 + 9 + j] = w[1] * Cx.m[3 + j];
        J[2 * 16 + 6 + j] = w[2] * RpwRi.m[6 + j];
        J[0 * 16 + 12 + j] = w[0] * Dq.m[j]; J[1 * 16 + 12 + j] = w[1] * Dq.m[3 + j];
        J[2 * 16 + 12 + j] = -w[2] * Eq.m[6 + j];
    }
    J[2 * 16 + 15] = w[2];
}

// ---------------------------------------------------------------------------------------------- manifolds
// PoseLocalParameterization / PoseSubsetParameterization::Plus (pose_local_parameterization.cpp:12-27,
// pose_subset_parameterization.cpp:27-53): p + dp, normalise(q * deltaQ(dtheta)); mask bit i zeroes delta[i].
VIWB_HD void pose_plus(const double *x, const double *delta, unsigned mask, double *out) {
    double d[6];
    for (int i = 0; i < 6; i++) d[i] = ((mask >> i) & 1u) ? 0.0 : delta[i];
    out[0] = x[0] + d[0]; out[1] = x[1] + d[1]; out[2] = x[2] + d[2];
    stq(out + 3, qnormalized(ldq(x + 3) * q_delta(v3(d[3], d[4], d[5]))));
}
// OrientationSubsetParameterization::Plus (orientation_subset_parameterization.cpp:27-44)
VIWB_HD void quat_plus(const double *x, const double *delta, unsigned mask, double *out) {
    double d[3];
    for (int i = 0; i < 3; i++) d[i] = ((mask >> i) & 1u) ? 0.0 : delta[i];
    stq(out, qnormalized(ldq(x) * q_delta(v3(d[0], d[1], d[2]))));
}

}  // namespace viwb

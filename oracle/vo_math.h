/*
 * vo_math.h -- small fixed-size linear algebra for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Restates the Eigen / Sophus / utility.h operations the reference's factors call, with the same
 * formulas and branch thresholds:
 *   - Eigen::Quaterniond product / inverse / toRotationMatrix / operator*(Vector3d)
 *   - Utility::deltaQ, skewSymmetric, Qleft, Qright, R2ypr, ypr2R   (utility/utility.h:22-113)
 *   - Sophus::SO3d::exp / log (Sophus a0fe89a, third party, restated from its published algorithm)
 *   - Sophus::rightJacobianSO3 / rightJacobianInvSO3                 (utility/sophus_utils.hpp:154-236)
 * Quaternions are stored [x,y,z,w] exactly like the reference's parameter blocks.
 * Matrices are row-major double arrays.
 */
#ifndef VO_MATH_H
#define VO_MATH_H
#include <math.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define VO_SOPHUS_EPS 1e-10       /* Sophus::Constants<double>::epsilon() */
#define VO_SOPHUS_EPS_SQRT 1e-5   /* Sophus::Constants<double>::epsilonSqrt() */

static inline void v3_set(double *o, double a, double b, double c) { o[0] = a; o[1] = b; o[2] = c; }
static inline void v3_copy(double *o, const double *a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static inline void v3_add(double *o, const double *a, const double *b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static inline void v3_sub(double *o, const double *a, const double *b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void v3_scale(double *o, const double *a, double s) { o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; }
static inline double v3_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double v3_norm(const double *a) { return sqrt(v3_dot(a, a)); }
static inline void v3_cross(double *o, const double *a, const double *b) {
    double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
    o[0] = t0; o[1] = t1; o[2] = t2;
}

/* 3x3 */
static inline void m3_identity(double *m) { memset(m, 0, 9 * sizeof(double)); m[0] = m[4] = m[8] = 1.0; }
static inline void m3_copy(double *o, const double *a) { memcpy(o, a, 9 * sizeof(double)); }
static inline void m3_transpose(double *o, const double *a) {
    double t[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[i * 3 + j] = a[j * 3 + i];
    memcpy(o, t, sizeof t);
}
static inline void m3_mul(double *o, const double *a, const double *b) {
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
        t[i * 3 + j] = s;
    }
    memcpy(o, t, sizeof t);
}
static inline void m3_mulv(double *o, const double *a, const double *v) {
    double t0 = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
    double t1 = a[3] * v[0] + a[4] * v[1] + a[5] * v[2];
    double t2 = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
static inline void m3_tmulv(double *o, const double *a, const double *v) { /* a^T v */
    double t0 = a[0] * v[0] + a[3] * v[1] + a[6] * v[2];
    double t1 = a[1] * v[0] + a[4] * v[1] + a[7] * v[2];
    double t2 = a[2] * v[0] + a[5] * v[1] + a[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
static inline void m3_scale(double *o, const double *a, double s) { for (int i = 0; i < 9; i++) o[i] = a[i] * s; }
static inline void m3_add(double *o, const double *a, const double *b) { for (int i = 0; i < 9; i++) o[i] = a[i] + b[i]; }
static inline void m3_sub(double *o, const double *a, const double *b) { for (int i = 0; i < 9; i++) o[i] = a[i] - b[i]; }
/* Utility::skewSymmetric (utility.h:38-46) */
static inline void m3_skew(double *o, const double *q) {
    o[0] = 0; o[1] = -q[2]; o[2] = q[1];
    o[3] = q[2]; o[4] = 0; o[5] = -q[0];
    o[6] = -q[1]; o[7] = q[0]; o[8] = 0;
}

/* generic dense row-major helpers */
static inline void mat_mul(double *o, const double *a, const double *b, int m, int k, int n) { /* o(mxn)=a(mxk) b(kxn); o must not alias */
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
        double s = 0; for (int l = 0; l < k; l++) s += a[i * k + l] * b[l * n + j];
        o[i * n + j] = s;
    }
}

/* ---- quaternions [x,y,z,w] ---------------------------------------------------------------- */
static inline void q_mul(double *o, const double *a, const double *b) { /* Eigen quaternion product a*b */
    double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
static inline void q_inv(double *o, const double *a) { /* Eigen::Quaternion::inverse(): conjugate / squaredNorm */
    double n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (n2 > 0) { o[0] = -a[0] / n2; o[1] = -a[1] / n2; o[2] = -a[2] / n2; o[3] = a[3] / n2; }
    else { o[0] = o[1] = o[2] = o[3] = 0; }
}
static inline void q_normalize(double *q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static inline void q_to_R(double *R, const double *q) { /* Eigen::QuaternionBase::toRotationMatrix */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static inline void q_rot(double *o, const double *q, const double *v) { /* Eigen q * v: v + w*uv + u x uv, uv = 2 u x v */
    double u[3] = {q[0], q[1], q[2]}, uv[3], uuv[3];
    v3_cross(uv, u, v); uv[0] *= 2; uv[1] *= 2; uv[2] *= 2;
    v3_cross(uuv, u, uv);
    double t0 = v[0] + q[3] * uv[0] + uuv[0], t1 = v[1] + q[3] * uv[1] + uuv[1], t2 = v[2] + q[3] * uv[2] + uuv[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
/* Eigen::Quaterniond(Matrix3d) (used by vector2double; estimator.cpp:1162) */
static inline void q_from_R(double *q, const double *m) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}
/* Utility::deltaQ (utility.h:22-36): normalise([1, theta/2]) */
static inline void q_delta(double *q, const double *theta) {
    q[0] = theta[0] / 2.0; q[1] = theta[1] / 2.0; q[2] = theta[2] / 2.0; q[3] = 1.0;
    q_normalize(q);
}
/* bottom-right 3x3 of Utility::Qleft / Qright (utility.h:58-76) */
static inline void q_left33(double *o, const double *q) {
    double s[9]; m3_skew(s, q);
    for (int i = 0; i < 9; i++) o[i] = s[i];
    o[0] += q[3]; o[4] += q[3]; o[8] += q[3];
}
static inline void q_right33(double *o, const double *q) {
    double s[9]; m3_skew(s, q);
    for (int i = 0; i < 9; i++) o[i] = -s[i];
    o[0] += q[3]; o[4] += q[3]; o[8] += q[3];
}

/* Utility::R2ypr / ypr2R in DEGREES (utility.h:78-113) */
static inline void R_to_ypr(double *ypr, const double *R) {
    double n[3] = {R[0], R[3], R[6]}, o[3] = {R[1], R[4], R[7]}, a[3] = {R[2], R[5], R[8]};
    double y = atan2(n[1], n[0]);
    double p = atan2(-n[2], n[0] * cos(y) + n[1] * sin(y));
    double r = atan2(a[0] * sin(y) - a[1] * cos(y), -o[0] * sin(y) + o[1] * cos(y));
    ypr[0] = y / M_PI * 180.0; ypr[1] = p / M_PI * 180.0; ypr[2] = r / M_PI * 180.0;
}
static inline void ypr_to_R(double *R, const double *ypr) {
    double y = ypr[0] / 180.0 * M_PI, p = ypr[1] / 180.0 * M_PI, r = ypr[2] / 180.0 * M_PI;
    double Rz[9] = {cos(y), -sin(y), 0, sin(y), cos(y), 0, 0, 0, 1};
    double Ry[9] = {cos(p), 0, sin(p), 0, 1, 0, -sin(p), 0, cos(p)};
    double Rx[9] = {1, 0, 0, 0, cos(r), -sin(r), 0, sin(r), cos(r)};
    double t[9]; m3_mul(t, Rz, Ry); m3_mul(R, t, Rx);
}

/* ---- Sophus SO3 (third party; restated) ------------------------------------------------------ */
static inline void so3_exp_q(double *q, const double *omega) { /* SO3d::exp(omega).unit_quaternion() */
    double theta_sq = v3_dot(omega, omega), theta = sqrt(theta_sq), half = 0.5 * theta, imag, real;
    if (theta < VO_SOPHUS_EPS) {
        double t4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
        real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * t4;
    } else {
        imag = sin(half) / theta; real = cos(half);
    }
    q[0] = imag * omega[0]; q[1] = imag * omega[1]; q[2] = imag * omega[2]; q[3] = real;
}
static inline void so3_exp_R(double *R, const double *omega) { double q[4]; so3_exp_q(q, omega); q_to_R(R, q); }
static inline void so3_log_q(double *omega, const double *q_in) { /* SO3d(q).log(): constructor normalises */
    double q[4] = {q_in[0], q_in[1], q_in[2], q_in[3]};
    q_normalize(q);
    double sq_n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], n = sqrt(sq_n), w = q[3], f;
    if (n < VO_SOPHUS_EPS) {
        double sq_w = w * w;
        f = 2.0 / w - 2.0 * sq_n / (w * sq_w);
    } else {
        if (fabs(w) < VO_SOPHUS_EPS) f = (w > 0 ? M_PI : -M_PI) / n;
        else f = 2.0 * atan(n / w) / n;
    }
    omega[0] = f * q[0]; omega[1] = f * q[1]; omega[2] = f * q[2];
}

/* Sophus::rightJacobianSO3 (sophus_utils.hpp:154-184) */
static inline void so3_Jr(double *J, const double *phi) {
    double n2 = v3_dot(phi, phi), h[9], h2[9];
    m3_skew(h, phi); m3_mul(h2, h, h); m3_identity(J);
    if (n2 > VO_SOPHUS_EPS) {
        double n = sqrt(n2), n3 = n2 * n, a = (1 - cos(n)) / n2, b = (n - sin(n)) / n3;
        for (int i = 0; i < 9; i++) J[i] += -h[i] * a + h2[i] * b;
    } else {
        for (int i = 0; i < 9; i++) J[i] += -h[i] / 2 + h2[i] / 6;
    }
}
/* Sophus::rightJacobianInvSO3 (sophus_utils.hpp:195-236) */
static inline void so3_Jr_inv(double *J, const double *phi) {
    double n2 = v3_dot(phi, phi), h[9], h2[9];
    m3_skew(h, phi); m3_mul(h2, h, h); m3_identity(J);
    for (int i = 0; i < 9; i++) J[i] += h[i] / 2;
    if (n2 > VO_SOPHUS_EPS) {
        double n = sqrt(n2);
        if (n < M_PI - VO_SOPHUS_EPS_SQRT) {
            double c = 1 / n2 - (1 + cos(n)) / (2 * n * sin(n));
            for (int i = 0; i < 9; i++) J[i] += h2[i] * c;
        } else {
            for (int i = 0; i < 9; i++) J[i] += h2[i] / (M_PI * M_PI);
        }
    } else {
        for (int i = 0; i < 9; i++) J[i] += h2[i] / 12;
    }
}
#endif

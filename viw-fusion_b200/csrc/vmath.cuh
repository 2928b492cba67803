// vmath.cuh -- small fixed-size FP64 math for the device code (vectors, 3x3, quaternions [x,y,z,w], SO(3)).
//
// Everything is VIWB_HD so the same source compiles for sm_100a (nvcc) and, under VIWB_HOST_EMU, for the
// CPU kernel-logic emulation used by the `not gpu` tests (tests/emu; never part of libviwb.so).
#pragma once
#include <math.h>

#ifdef VIWB_HOST_EMU
#define VIWB_HD static inline
#define VIWB_D static inline
#define VIWB_DM inline
#define VIWB_SYNC() ((void)0)
#define VIWB_SYNCWARP() ((void)0)
#define VIWB_RESTRICT
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }      // CUDA's device intrinsic, for the host emulation
#else
#define VIWB_HD __host__ __device__ __forceinline__
#define VIWB_D __device__ __forceinline__
#define VIWB_DM __device__ __forceinline__
#define VIWB_SYNC() __syncthreads()
#define VIWB_SYNCWARP() __syncwarp()
#define VIWB_RESTRICT __restrict__
#endif

namespace viwb {

struct V3 { double x, y, z; };
struct M3 { double m[9]; };   // row-major
struct Q4 { double x, y, z, w; };

VIWB_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
VIWB_HD V3 operator+(const V3 &a, const V3 &b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
VIWB_HD V3 operator-(const V3 &a, const V3 &b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
VIWB_HD V3 operator-(const V3 &a) { return v3(-a.x, -a.y, -a.z); }
VIWB_HD V3 operator*(const V3 &a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
VIWB_HD V3 operator*(double s, const V3 &a) { return v3(a.x * s, a.y * s, a.z * s); }
VIWB_HD double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VIWB_HD V3 cross(const V3 &a, const V3 &b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
VIWB_HD V3 ld3(const double *p) { return v3(p[0], p[1], p[2]); }
VIWB_HD void st3(double *p, const V3 &a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
VIWB_HD double comp(const V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

VIWB_HD M3 m3_identity() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = 0.0; r.m[0] = r.m[4] = r.m[8] = 1.0; return r; }
VIWB_HD M3 m3_ld(const double *p) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = p[i]; return r; }
VIWB_HD M3 operator*(const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
    return r;
}
VIWB_HD V3 operator*(const M3 &a, const V3 &v) {
    return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
VIWB_HD M3 operator*(const M3 &a, double s) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] * s; return r; }
VIWB_HD M3 operator+(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i]; return r; }
VIWB_HD M3 operator-(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i]; return r; }
VIWB_HD M3 operator-(const M3 &a) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = -a.m[i]; return r; }
VIWB_HD M3 transpose(const M3 &a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[j * 3 + i]; return r; }
VIWB_HD V3 tmul(const M3 &a, const V3 &v) {   // a^T v
    return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
VIWB_HD M3 skew(const V3 &q) { M3 r; r.m[0] = 0; r.m[1] = -q.z; r.m[2] = q.y; r.m[3] = q.z; r.m[4] = 0; r.m[5] = -q.x; r.m[6] = -q.y; r.m[7] = q.x; r.m[8] = 0; return r; }
VIWB_HD M3 outer(const V3 &a, const V3 &b) { M3 r; r.m[0] = a.x * b.x; r.m[1] = a.x * b.y; r.m[2] = a.x * b.z; r.m[3] = a.y * b.x; r.m[4] = a.y * b.y; r.m[5] = a.y * b.z; r.m[6] = a.z * b.x; r.m[7] = a.z * b.y; r.m[8] = a.z * b.z; return r; }
VIWB_HD V3 row(const M3 &a, int i) { return v3(a.m[i * 3], a.m[i * 3 + 1], a.m[i * 3 + 2]); }

// quaternions
VIWB_HD Q4 q4(double x, double y, double z, double w) { Q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
VIWB_HD Q4 ldq(const double *p) { return q4(p[0], p[1], p[2], p[3]); }
VIWB_HD void stq(double *p, const Q4 &q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
VIWB_HD V3 qvec(const Q4 &q) { return v3(q.x, q.y, q.z); }
VIWB_HD Q4 operator*(const Q4 &a, const Q4 &b) {
    return q4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
VIWB_HD Q4 qinv(const Q4 &a) {   // conjugate / squared norm (Eigen::Quaternion::inverse)
    double n2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w, s = 1.0 / n2;
    return q4(-a.x * s, -a.y * s, -a.z * s, a.w * s);
}
VIWB_HD Q4 qnormalized(const Q4 &a) {
    double s = 1.0 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
    return q4(a.x * s, a.y * s, a.z * s, a.w * s);
}
VIWB_HD M3 qR(const Q4 &q) {
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 r;
    r.m[0] = 1 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy;
    r.m[3] = txy + twz; r.m[4] = 1 - (txx + tzz); r.m[5] = tyz - twx;
    r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1 - (txx + tyy);
    return r;
}
VIWB_HD V3 qrot(const Q4 &q, const V3 &v) { V3 u = qvec(q); V3 uv = 2.0 * cross(u, v); return v + q.w * uv + cross(u, uv); }
VIWB_HD Q4 q_from_R(const M3 &a) {   // Eigen::Quaterniond(Matrix3d)
    const double *m = a.m; double t = m[0] + m[4] + m[8]; double q[4];
    if (t > 0) { t = sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t; q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t; }
    else {
        int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0); q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t; q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t; q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
    return q4(q[0], q[1], q[2], q[3]);
}
// Utility::deltaQ (utility.h:22-36): normalise([1, theta/2])
VIWB_HD Q4 q_delta(const V3 &th) { return qnormalized(q4(th.x * 0.5, th.y * 0.5, th.z * 0.5, 1.0)); }
// bottom-right 3x3 of Utility::Qleft / Qright (utility.h:58-76)
VIWB_HD M3 q_left33(const Q4 &q) { M3 r = skew(qvec(q)); r.m[0] += q.w; r.m[4] += q.w; r.m[8] += q.w; return r; }
VIWB_HD M3 q_right33(const Q4 &q) { M3 r = -skew(qvec(q)); r.m[0] += q.w; r.m[4] += q.w; r.m[8] += q.w; return r; }

// Sophus SO3 (restated): exp to quaternion, log of a quaternion, right Jacobian and its inverse
#define VIWB_SOPHUS_EPS 1e-10
#define VIWB_SOPHUS_EPS_SQRT 1e-5
VIWB_HD Q4 so3_exp_q(const V3 &w) {
    double t2 = dot(w, w), t = sqrt(t2), im, re;
    if (t < VIWB_SOPHUS_EPS) { double t4 = t2 * t2; im = 0.5 - t2 / 48.0 + t4 / 3840.0; re = 1.0 - t2 / 8.0 + t4 / 384.0; }
    else { double h = 0.5 * t; im = sin(h) / t; re = cos(h); }
    return q4(im * w.x, im * w.y, im * w.z, re);
}
VIWB_HD M3 so3_exp_R(const V3 &w) { return qR(so3_exp_q(w)); }
VIWB_HD V3 so3_log_q(const Q4 &qi) {
    Q4 q = qnormalized(qi);
    double n2 = q.x * q.x + q.y * q.y + q.z * q.z, n = sqrt(n2), w = q.w, f;
    if (n < VIWB_SOPHUS_EPS) f = 2.0 / w - 2.0 * n2 / (w * w * w);
    else if (fabs(w) < VIWB_SOPHUS_EPS) f = (w > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    else f = 2.0 * atan(n / w) / n;
    return v3(f * q.x, f * q.y, f * q.z);
}
VIWB_HD M3 so3_Jr(const V3 &phi) {
    double n2 = dot(phi, phi); M3 h = skew(phi), h2 = h * h, J = m3_identity();
    if (n2 > VIWB_SOPHUS_EPS) { double n = sqrt(n2); J = J - h * ((1 - cos(n)) / n2) + h2 * ((n - sin(n)) / (n2 * n)); }
    else J = J - h * 0.5 + h2 * (1.0 / 6.0);
    return J;
}
VIWB_HD M3 so3_Jr_inv(const V3 &phi) {
    double n2 = dot(phi, phi); M3 h = skew(phi), h2 = h * h, J = m3_identity() + h * 0.5;
    if (n2 > VIWB_SOPHUS_EPS) {
        double n = sqrt(n2);
        if (n < 3.14159265358979323846 - VIWB_SOPHUS_EPS_SQRT) J = J + h2 * (1.0 / n2 - (1 + cos(n)) / (2 * n * sin(n)));
        else J = J + h2 * (1.0 / (3.14159265358979323846 * 3.14159265358979323846));
    } else J = J + h2 * (1.0 / 12.0);
    return J;
}
// Utility::R2ypr / ypr2R (degrees, utility.h:78-113)
VIWB_HD V3 R_to_ypr(const M3 &R) {
    double y = atan2(R.m[3], R.m[0]);
    double p = atan2(-R.m[6], R.m[0] * cos(y) + R.m[3] * sin(y));
    double r = atan2(R.m[2] * sin(y) - R.m[5] * cos(y), -R.m[1] * sin(y) + R.m[4] * cos(y));
    const double k = 180.0 / 3.14159265358979323846;
    return v3(y * k, p * k, r * k);
}
VIWB_HD M3 yaw_to_R(double yaw_deg) {
    double y = yaw_deg / 180.0 * 3.14159265358979323846; M3 r = m3_identity();
    r.m[0] = cos(y); r.m[1] = -sin(y); r.m[3] = sin(y); r.m[4] = cos(y);
    return r;
}

}  // namespace viwb

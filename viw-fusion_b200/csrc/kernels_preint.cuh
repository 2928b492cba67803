// kernels_preint.cuh -- IMU / wheel pre-integration (the producers of the constants that a-8 / a-9 consume):
//   IntegrationBase::midPointIntegration / propagate        factor/integration_base.h:63-167
//   WheelIntegrationBase::midPointIntegration / propagate   factor/wheel_integration_base.h:67-177
// One warp per interval (the samples between two key-frames, 10-40 steps): lane 0 advances the state and builds the step's
// F (15x15 / 6x6) and V (15x18 / 6x12) in the warp's shared-memory slice, the lanes share the covariance / Jacobian products
// jacobian <- F jacobian, covariance <- F cov F^T + V Q V^T.  Steps are sequential by nature; intervals are independent.
#pragma once
#include "factors.cuh"

namespace viwb {

struct ImuPreArgs { int n; const int *off; const double *dt, *acc, *gyr, *ba, *bg; double noise[4]; double *rec; };
struct WheelPreArgs { int n; const int *off; const double *dt, *vel, *gyr, *s, *td; double noise[2]; double *rec; };
enum { PRE_IMU_SMEM = 225 * 5 + 270, PRE_WHEEL_SMEM = 36 * 4 + 72 + 18 };      // doubles per warp


VIWB_D void imu_preint_warp(const ImuPreArgs &a, int it, int lane, int W, double *sm) {
    if (it >= a.n) return;
    double *jac = sm, *cov = jac + 225, *F = cov + 225, *V = F + 225, *T1 = V + 270, *T2 = T1 + 225;
    const int s0 = a.off[it], cnt = a.off[it + 1] - s0, r0 = s0 + it;      // (cnt + 1) sample rows per interval
    const V3 ba = ld3(a.ba + 3 * it), bg = ld3(a.bg + 3 * it);
    for (int e = lane; e < 225; e += W) { jac[e] = (e / 15 == e % 15) ? 1.0 : 0.0; cov[e] = 0.0; }
    V3 dp = v3(0, 0, 0), dv = v3(0, 0, 0); Q4 dq = q4(0, 0, 0, 1);
    double sum_dt = 0.0;
    const double an2 = a.noise[0] * a.noise[0], gn2 = a.noise[1] * a.noise[1], aw2 = a.noise[2] * a.noise[2], gw2 = a.noise[3] * a.noise[3];
    VIWB_SYNCWARP();
    for (int s = 0; s < cnt; s++) {
        const double dt = a.dt[s0 + s];
        const V3 a0 = ld3(a.acc + 3 * (r0 + s)), g0 = ld3(a.gyr + 3 * (r0 + s)), a1 = ld3(a.acc + 3 * (r0 + s + 1)), g1 = ld3(a.gyr + 3 * (r0 + s + 1));
        // midPointIntegration (every lane carries the small state redundantly; lane 0 writes F and V)
        const V3 un_acc_0 = qrot(dq, a0 - ba);
        const V3 un_gyr = 0.5 * (g0 + g1) - bg;
        const Q4 rq = dq * q4(un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2, 1.0);
        const V3 un_acc_1 = qrot(rq, a1 - ba);
        const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
        const V3 rp = dp + dv * dt + 0.5 * un_acc * dt * dt, rv = dv + un_acc * dt;
        if (lane == 0) {
            const M3 R_w_x = skew(un_gyr), R_a_0_x = skew(a0 - ba), R_a_1_x = skew(a1 - ba);
            const M3 Rd = qR(dq), Rr = qR(rq), I = m3_identity();
            const M3 IwR = I - R_w_x * dt, Rd_Ra0 = Rd * R_a_0_x, Rr_Ra1 = Rr * R_a_1_x, Rr_Ra1_IwR = Rr_Ra1 * IwR;
            for (int e = 0; e < 225; e++) F[e] = 0.0;
            for (int e = 0; e < 270; e++) V[e] = 0.0;
            put33(F, 15, 0, 0, I);
            put33(F, 15, 0, 3, Rd_Ra0 * (-0.25 * dt * dt) + Rr_Ra1_IwR * (-0.25 * dt * dt));
            put33(F, 15, 0, 6, I * dt);
            put33(F, 15, 0, 9, (Rd + Rr) * (-0.25 * dt * dt));
            put33(F, 15, 0, 12, Rr_Ra1 * (-0.25 * dt * dt * -dt));
            put33(F, 15, 3, 3, IwR);
            put33(F, 15, 3, 12, I * (-1.0 * dt));
            put33(F, 15, 6, 3, Rd_Ra0 * (-0.5 * dt) + Rr_Ra1_IwR * (-0.5 * dt));
            put33(F, 15, 6, 6, I);
            put33(F, 15, 6, 9, (Rd + Rr) * (-0.5 * dt));
            put33(F, 15, 6, 12, Rr_Ra1 * (-0.5 * dt * -dt));
            put33(F, 15, 9, 9, I); put33(F, 15, 12, 12, I);
            put33(V, 18, 0, 0, Rd * (0.25 * dt * dt));
            const M3 v03 = (-Rr_Ra1) * (0.25 * dt * dt * 0.5 * dt);
            put33(V, 18, 0, 3, v03); put33(V, 18, 0, 9, v03);
            put33(V, 18, 0, 6, Rr * (0.25 * dt * dt));
            put33(V, 18, 3, 3, I * (0.5 * dt)); put33(V, 18, 3, 9, I * (0.5 * dt));
            put33(V, 18, 6, 0, Rd * (0.5 * dt));
            const M3 v63 = (-Rr_Ra1) * (0.5 * dt * 0.5 * dt);
            put33(V, 18, 6, 3, v63); put33(V, 18, 6, 9, v63);
            put33(V, 18, 6, 6, Rr * (0.5 * dt));
            put33(V, 18, 9, 12, I * dt); put33(V, 18, 12, 15, I * dt);
        }
        VIWB_SYNCWARP();
        for (int e = lane; e < 225; e += W) {          // T1 = F jac ; T2 = F cov
            const int i = e / 15, j = e % 15;
            double s1 = 0.0, s2 = 0.0;
            for (int k = 0; k < 15; k++) { s1 += F[i * 15 + k] * jac[k * 15 + j]; s2 += F[i * 15 + k] * cov[k * 15 + j]; }
            T1[e] = s1; T2[e] = s2;
        }
        VIWB_SYNCWARP();
        for (int e = lane; e < 225; e += W) {          // cov = T2 F^T + V Q V^T ; jac = T1
            const int i = e / 15, j = e % 15;
            double sF = 0.0, sV = 0.0;
            for (int k = 0; k < 15; k++) sF += T2[i * 15 + k] * F[j * 15 + k];
            for (int k = 0; k < 18; k++) { const double q = (k < 3 || (k >= 6 && k < 9)) ? an2 : (k < 12 ? gn2 : (k < 15 ? aw2 : gw2)); sV += (V[i * 18 + k] * q) * V[j * 18 + k]; }
            cov[e] = sF + sV; jac[e] = T1[e];
        }
        VIWB_SYNCWARP();
        dp = rp; dv = rv; dq = qnormalized(rq);
        sum_dt += dt;
    }
    double *rec = a.rec + (size_t)it * 287;
    if (lane == 0) {
        rec[0] = sum_dt; st3(rec + 1, dp); stq(rec + 4, dq); st3(rec + 8, dv); st3(rec + 11, ba); st3(rec + 14, bg);
        const int blk[5][2] = {{0, 9}, {0, 12}, {3, 12}, {6, 9}, {6, 12}};   // dp_dba dp_dbg dq_dbg dv_dba dv_dbg
        for (int b = 0; b < 5; b++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rec[17 + 9 * b + i * 3 + j] = jac[(blk[b][0] + i) * 15 + blk[b][1] + j];
    }
    for (int e = lane; e < 225; e += W) rec[62 + e] = cov[e];
}

VIWB_D void wheel_preint_warp(const WheelPreArgs &a, int it, int lane, int W, double *sm) {
    if (it >= a.n) return;
    double *cov = sm, *F = cov + 36, *V = F + 36, *T1 = V + 72, *T2 = T1 + 36, *jac = T2 + 36;       // jac 6x3
    const int s0 = a.off[it], cnt = a.off[it + 1] - s0, r0 = s0 + it;
    const double sx = a.s[3 * it], sy = a.s[3 * it + 1], sw = a.s[3 * it + 2];
    const V3 sv = v3(sx, sy, 1.0);
    for (int e = lane; e < 36; e += W) cov[e] = 0.0;
    for (int e = lane; e < 18; e += W) jac[e] = 0.0;
    V3 dp = v3(0, 0, 0); Q4 dq = q4(0, 0, 0, 1);
    double sum_dt = 0.0;
    const double vn2 = a.noise[0] * a.noise[0], gn2 = a.noise[1] * a.noise[1];
    VIWB_SYNCWARP();
    for (int s = 0; s < cnt; s++) {
        const double dt = a.dt[s0 + s];
        const V3 v0 = ld3(a.vel + 3 * (r0 + s)), g0 = ld3(a.gyr + 3 * (r0 + s)), v1 = ld3(a.vel + 3 * (r0 + s + 1)), g1 = ld3(a.gyr + 3 * (r0 + s + 1));
        const V3 sv_v0 = v3(sv.x * v0.x, sv.y * v0.y, sv.z * v0.z), sv_v1 = v3(sv.x * v1.x, sv.y * v1.y, sv.z * v1.z);
        const V3 un_vel_0 = qrot(dq, sv_v0);
        const V3 un_gyr = (0.5 * sw) * (g0 + g1);
        const Q4 ddq = q4(un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2, 1.0);
        const Q4 rq = dq * ddq;
        const V3 un_vel_1 = qrot(rq, sv_v1);
        const V3 rp = dp + 0.5 * (un_vel_0 + un_vel_1) * dt;
        if (lane == 0) {
            const M3 R_v0 = skew(sv_v0), R_v1 = skew(sv_v1), Rd = qR(dq), Rr = qR(rq), RddT = transpose(qR(ddq));
            for (int e = 0; e < 36; e++) F[e] = 0.0;
            for (int e = 0; e < 72; e++) V[e] = 0.0;
            put33(F, 6, 0, 0, m3_identity());
            put33(F, 6, 0, 3, (Rd * R_v0 + (Rr * R_v1) * RddT) * (-0.5 * dt));
            put33(F, 6, 3, 3, RddT);
            const M3 Jr = so3_Jr(un_gyr * dt);
            M3 m; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m.m[i * 3 + j] = 0.5 * dt * Rd.m[i * 3 + j] * comp(sv, j);
            put33(V, 12, 0, 0, m);
            const M3 m2 = ((Rr * R_v1) * Jr) * (-0.25 * dt * dt);
            put33(V, 12, 0, 3, m2); put33(V, 12, 0, 9, m2);
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m.m[i * 3 + j] = 0.5 * dt * Rr.m[i * 3 + j] * comp(sv, j);
            put33(V, 12, 0, 6, m);
            put33(V, 12, 3, 3, Jr * (0.5 * sw * dt)); put33(V, 12, 3, 9, Jr * (0.5 * sw * dt));
            // Jacobian w.r.t. the intrinsics (sx, sy, sw)
            for (int k = 0; k < 3; k++) {
                jac[k * 3 + 0] += 0.5 * (Rd.m[k * 3 + 0] * v0.x + Rr.m[k * 3 + 0] * v1.x) * dt;
                jac[k * 3 + 1] += 0.5 * (Rd.m[k * 3 + 1] * v0.y + Rr.m[k * 3 + 1] * v1.y) * dt;
            }
            const V3 dr_last = v3(jac[9 + 2], jac[12 + 2], jac[15 + 2]);
            const V3 Jg = Jr * ((0.5 * (g0 + g1)) * dt);
            jac[9 + 2] += Jg.x; jac[12 + 2] += Jg.y; jac[15 + 2] += Jg.z;
            const V3 dr_new = v3(jac[9 + 2], jac[12 + 2], jac[15 + 2]);
            const V3 ta = Rd * (skew(dr_last) * sv_v0), tb = Rr * (skew(dr_new) * sv_v1);
            jac[0 * 3 + 2] += 0.5 * (ta.x + tb.x) * dt; jac[1 * 3 + 2] += 0.5 * (ta.y + tb.y) * dt; jac[2 * 3 + 2] += 0.5 * (ta.z + tb.z) * dt;
        }
        VIWB_SYNCWARP();
        for (int e = lane; e < 36; e += W) { const int i = e / 6, j = e % 6; double s1 = 0.0; for (int k = 0; k < 6; k++) s1 += F[i * 6 + k] * cov[k * 6 + j]; T1[e] = s1; }
        VIWB_SYNCWARP();
        for (int e = lane; e < 36; e += W) {
            const int i = e / 6, j = e % 6;
            double sF = 0.0, sV = 0.0;
            for (int k = 0; k < 6; k++) sF += T1[i * 6 + k] * F[j * 6 + k];
            for (int k = 0; k < 12; k++) { const double q = (k < 3 || (k >= 6 && k < 9)) ? vn2 : gn2; sV += (V[i * 12 + k] * q) * V[j * 12 + k]; }
            T2[e] = sF + sV;
        }
        VIWB_SYNCWARP();
        for (int e = lane; e < 36; e += W) cov[e] = T2[e];
        VIWB_SYNCWARP();
        dp = rp; dq = qnormalized(rq);
        sum_dt += dt;
    }
    double *rec = a.rec + (size_t)it * 78;
    if (lane == 0) {
        st3(rec, dp); stq(rec + 3, dq);
        for (int e = 0; e < 18; e++) rec[7 + e] = jac[e];
        rec[61] = sx; rec[62] = sy; rec[63] = sw; rec[64] = a.td[it];
        for (int k = 0; k < 3; k++) { rec[65 + k] = a.vel[3 * r0 + k]; rec[68 + k] = a.gyr[3 * r0 + k]; rec[71 + k] = a.vel[3 * (r0 + cnt) + k]; rec[74 + k] = a.gyr[3 * (r0 + cnt) + k]; }
        rec[77] = sum_dt;
    }
    for (int e = lane; e < 36; e += W) rec[25 + e] = cov[e];
}

#ifndef VIWB_HOST_EMU
enum { PRE_WPB = 4 };
__global__ void __launch_bounds__(32 * PRE_WPB) imu_preint_kernel(ImuPreArgs a) {
    extern __shared__ double pre_smem[];
    const int warp = threadIdx.x >> 5;
    imu_preint_warp(a, blockIdx.x * PRE_WPB + warp, threadIdx.x & 31, 32, pre_smem + (size_t)warp * PRE_IMU_SMEM);
}
__global__ void __launch_bounds__(32 * PRE_WPB) wheel_preint_kernel(WheelPreArgs a) {
    extern __shared__ double pre_smem[];
    const int warp = threadIdx.x >> 5;
    wheel_preint_warp(a, blockIdx.x * PRE_WPB + warp, threadIdx.x & 31, 32, pre_smem + (size_t)warp * PRE_WHEEL_SMEM);
}
#endif

}  // namespace viwb

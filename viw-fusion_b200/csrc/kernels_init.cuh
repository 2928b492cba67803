// kernels_init.cuh -- visual-inertial(-wheel) alignment of the initialisation (SURVEY 8 f-4 ii; runs once per session, before the first window solve):
//   solveGyroscopeBias                       initial/initial_aligment.cpp:14-48
//   TangentBasis                             :51-64
//   LinearAlignment / RefineGravity          :66-203   (camera + IMU)
//   LinearAlignmentWithWheel / RefineGravityWithWheel   :204-334   (adds the wheel odometer's delta_p rows)
// One block per call.  The normal equations are at most (3*64+4)^2 doubles and live in a global-memory workspace (L2-resident); thread 0
// accumulates the per-interval 10x10 / 9x9 contributions in the reference's order (they overlap between neighbouring intervals), the block
// shares the pivoted LDL^T factorisation and the triangular solves.  Reference behaviours kept on purpose: the scale unknown is carried as
// 100*s (the /100.0 in the last column), both systems are multiplied by 1000 before the solve, and RefineGravity does NOT clear A and b
// between its four iterations -- each iteration adds its contributions to 1000x the previous system (:138-140, :263-265).
#pragma once
#include "factors.cuh"

namespace viwb {

struct AlignArgs {
    int F, use_wheel;
    const double *R, *T;        // [F*9] row-major ImageFrame::R, [F*3] ImageFrame::T
    const double *imu;          // [(F-1)*VIWB_IMU_DOUBLES]   pre_integration of frame j = i+1 (record i)
    const double *wheel;        // [(F-1)*VIWB_WHEEL_DOUBLES] or nullptr
    double tic[3], rio[9], tio[3], g_norm;
    double *A, *b, *x;          // workspace: n*n, n, n   (n = 3F+4)
    int *perm;                  // workspace: n
    double *out;                // [0] return value (1 / 0), [1..3] g, [4] n_x, [5 ...] x (3F+4 after a failed first stage, else 3F+3 with x.tail = s)
};
struct GyroBiasArgs { int F; const double *R, *imu; double *A, *b, *x; int *perm; double *out; };   // out[0..2] = delta_bg

// x = A.ldlt().solve(b): LDL^T with diagonal pivoting (largest remaining |diagonal|, as Eigen's LDLT), in place on the lower triangle of A.
VIWB_D void ldlt_solve_block(double *A, int n, const double *b, double *x, int *perm, int tid, int nt) {
    for (int k = 0; k < n; k++) {
        if (tid == 0) {
            int p = k; double best = fabs(A[k * n + k]);
            for (int i = k + 1; i < n; i++) { const double v = fabs(A[i * n + i]); if (v > best) { best = v; p = i; } }
            perm[k] = p;
        }
        VIWB_SYNC();
        const int p = perm[k];
        if (p != k) {        // symmetric exchange of rows / columns k and p, lower triangle only
            for (int j = tid; j < k; j += nt) { const double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for (int i = p + 1 + tid; i < n; i += nt) { const double t = A[i * n + k]; A[i * n + k] = A[i * n + p]; A[i * n + p] = t; }
            for (int i = k + 1 + tid; i < p; i += nt) { const double t = A[i * n + k]; A[i * n + k] = A[p * n + i]; A[p * n + i] = t; }
            if (tid == 0) { const double t = A[k * n + k]; A[k * n + k] = A[p * n + p]; A[p * n + p] = t; }
            VIWB_SYNC();
        }
        const double d = A[k * n + k];
        const double inv = fabs(d) > 2.2250738585072014e-308 ? 1.0 / d : 0.0;
        for (int i = k + 1 + tid; i < n; i += nt) {          // trailing update with the unscaled column, row by row
            const double lik = A[i * n + k] * inv;
            for (int j = k + 1; j <= i; j++) A[i * n + j] -= lik * A[j * n + k];
        }
        VIWB_SYNC();
        for (int i = k + 1 + tid; i < n; i += nt) A[i * n + k] *= inv;
        VIWB_SYNC();
    }
    // solve: x = P^T L^-T D^+ L^-1 P b   (sequential substitutions; n <= 196)
    if (tid == 0) {
        for (int i = 0; i < n; i++) x[i] = b[i];
        for (int k = 0; k < n; k++) { const int p = perm[k]; if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; } }
        for (int i = 0; i < n; i++) { double s = x[i]; for (int j = 0; j < i; j++) s -= A[i * n + j] * x[j]; x[i] = s; }
        for (int i = 0; i < n; i++) { const double d = A[i * n + i]; x[i] = fabs(d) > 2.2250738585072014e-308 ? x[i] / d : 0.0; }
        for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int j = i + 1; j < n; j++) s -= A[j * n + i] * x[j]; x[i] = s; }
        for (int k = n - 1; k >= 0; k--) { const int p = perm[k]; if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; } }
    }
    VIWB_SYNC();
}

VIWB_D void gyro_bias_block(const GyroBiasArgs &a, int tid, int nt) {
    if (tid == 0) {
        double A[9], b[3];
        for (int e = 0; e < 9; e++) A[e] = 0.0;
        b[0] = b[1] = b[2] = 0.0;
        for (int i = 0; i + 1 < a.F; i++) {
            const M3 Ri = m3_ld(a.R + 9 * i), Rj = m3_ld(a.R + 9 * (i + 1));
            const double *c = a.imu + (size_t)i * VIWB_IMU_DOUBLES;
            const Q4 q_ij = q_from_R(transpose(Ri) * Rj);
            const M3 J = m3_ld(c + 35);                                          // jacobian.block<3,3>(O_R, O_BG)
            const V3 tb = 2.0 * qvec(qinv(ldq(c + 4)) * q_ij);
            const M3 JtJ = transpose(J) * J; const V3 Jtb = tmul(J, tb);
            for (int e = 0; e < 9; e++) A[e] += JtJ.m[e];
            b[0] += Jtb.x; b[1] += Jtb.y; b[2] += Jtb.z;
        }
        for (int e = 0; e < 9; e++) a.A[e] = A[e];
        for (int e = 0; e < 3; e++) a.b[e] = b[e];
    }
    VIWB_SYNC();
    ldlt_solve_block(a.A, 3, a.b, a.x, a.perm, tid, nt);
    if (tid == 0) for (int e = 0; e < 3; e++) a.out[e] = a.x[e];
}

// tmp_A (rows x cols, row-major in `ta`) and tmp_b of interval i; refine = the 9-column form with the tangent basis lxly and the g0 terms
VIWB_D void align_rows(const AlignArgs &a, int i, bool refine, const V3 &g0, const V3 &lx, const V3 &ly, double *ta, double *tb, int rows, int cols) {
    const M3 Ri = m3_ld(a.R + 9 * i), Rj = m3_ld(a.R + 9 * (i + 1)), RiT = transpose(Ri);
    const V3 Ti = ld3(a.T + 3 * i), Tj = ld3(a.T + 3 * (i + 1)), tic = ld3(a.tic);
    const double *c = a.imu + (size_t)i * VIWB_IMU_DOUBLES;
    const double dt = c[0];
    const V3 dp = ld3(c + 1), dv = ld3(c + 8);
    for (int e = 0; e < rows * cols; e++) ta[e] = 0.0;
    for (int e = 0; e < rows; e++) tb[e] = 0.0;
    const M3 RiTRj = RiT * Rj;
    const M3 half = RiT * (dt * dt / 2), full = RiT * dt;
    for (int r = 0; r < 3; r++) {
        ta[r * cols + r] = -dt;
        ta[(3 + r) * cols + r] = -1.0;
        for (int q = 0; q < 3; q++) ta[(3 + r) * cols + 3 + q] = RiTRj.m[r * 3 + q];
    }
    const V3 sc = (RiT * (Tj - Ti)) * (1.0 / 100.0);
    const int cs = cols - 1;                                                     // the scale column
    ta[0 * cols + cs] = sc.x; ta[1 * cols + cs] = sc.y; ta[2 * cols + cs] = sc.z;
    V3 b0 = dp + RiTRj * tic - tic, b1 = dv;
    if (!refine) {
        for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) { ta[r * cols + 6 + q] = half.m[r * 3 + q]; ta[(3 + r) * cols + 6 + q] = full.m[r * 3 + q]; }
    } else {
        const V3 hx = half * lx, hy = half * ly, fx = full * lx, fy = full * ly;
        ta[0 * cols + 6] = hx.x; ta[1 * cols + 6] = hx.y; ta[2 * cols + 6] = hx.z; ta[0 * cols + 7] = hy.x; ta[1 * cols + 7] = hy.y; ta[2 * cols + 7] = hy.z;
        ta[3 * cols + 6] = fx.x; ta[4 * cols + 6] = fx.y; ta[5 * cols + 6] = fx.z; ta[3 * cols + 7] = fy.x; ta[4 * cols + 7] = fy.y; ta[5 * cols + 7] = fy.z;
        b0 = b0 - half * g0; b1 = b1 - full * g0;
    }
    tb[0] = b0.x; tb[1] = b0.y; tb[2] = b0.z; tb[3] = b1.x; tb[4] = b1.y; tb[5] = b1.z;
    if (rows == 9) {                                                             // the wheel odometer rows (:235-236, :301-302)
        const M3 rio = m3_ld(a.rio), rioT = transpose(rio); const V3 tio = ld3(a.tio);
        const M3 RiRioT = transpose(Ri * rio);
        const V3 sw = (RiRioT * (Tj - Ti)) * (1.0 / 100.0);
        ta[6 * cols + cs] = sw.x; ta[7 * cols + cs] = sw.y; ta[8 * cols + cs] = sw.z;
        const V3 wdp = ld3(a.wheel + (size_t)i * VIWB_WHEEL_DOUBLES);
        const V3 b2 = wdp - rioT * (RiTRj * tio) + RiRioT * (Rj * tic) - rioT * (tic - tio);
        tb[6] = b2.x; tb[7] = b2.y; tb[8] = b2.z;
    }
}

// adds r_A = tmp_A^T tmp_A and r_b = tmp_A^T tmp_b of interval i into the big system: the 6x6 block at (3i,3i), the `tailn` x `tailn` corner, and the two cross blocks
VIWB_D void align_accumulate(double *A, double *b, int n, int i, const double *ta, const double *tb, int rows, int cols) {
    const int tailn = cols - 6;
    for (int p = 0; p < cols; p++) {
        const int gp = p < 6 ? 3 * i + p : n - tailn + (p - 6);
        double sb = 0.0;
        for (int r = 0; r < rows; r++) sb += ta[r * cols + p] * tb[r];
        b[gp] += sb;
        for (int q = 0; q < cols; q++) {
            const int gq = q < 6 ? 3 * i + q : n - tailn + (q - 6);
            double s = 0.0;
            for (int r = 0; r < rows; r++) s += ta[r * cols + p] * ta[r * cols + q];
            A[gp * n + gq] += s;
        }
    }
}

VIWB_D void align_block(const AlignArgs &a, int tid, int nt) {
    const int F = a.F, rows = a.use_wheel ? 9 : 6;
    double ta[90], tb[9];
    // ---- LinearAlignment(WithWheel): unknowns = F velocities (3 each), g (3), 100*s
    int n = 3 * F + 4;
    for (int e = tid; e < n * n; e += nt) a.A[e] = 0.0;
    for (int e = tid; e < n; e += nt) a.b[e] = 0.0;
    VIWB_SYNC();
    if (tid == 0) {
        const V3 z = v3(0, 0, 0);
        for (int i = 0; i + 1 < F; i++) { align_rows(a, i, false, z, z, z, ta, tb, rows, 10); align_accumulate(a.A, a.b, n, i, ta, tb, rows, 10); }
    }
    VIWB_SYNC();
    for (int e = tid; e < n * n; e += nt) a.A[e] *= 1000.0;
    for (int e = tid; e < n; e += nt) a.b[e] *= 1000.0;
    VIWB_SYNC();
    ldlt_solve_block(a.A, n, a.b, a.x, a.perm, tid, nt);
    const double s1 = a.x[n - 1] / 100.0;
    V3 g = v3(a.x[n - 4], a.x[n - 3], a.x[n - 2]);
    const double gn = sqrt(dot(g, g));
    if (fabs(gn - a.g_norm) > 0.5 || s1 < 0) {
        if (tid == 0) { a.out[0] = 0.0; st3(a.out + 1, g); a.out[4] = (double)n; }
        for (int e = tid; e < n; e += nt) a.out[5 + e] = a.x[e];
        return;
    }
    // ---- RefineGravity(WithWheel): g on its tangent plane, 4 passes, A and b carried over (x1000) from pass to pass as the reference does
    n = 3 * F + 3;
    VIWB_SYNC();
    for (int e = tid; e < n * n; e += nt) a.A[e] = 0.0;
    for (int e = tid; e < n; e += nt) a.b[e] = 0.0;
    V3 g0 = g * (1.0 / gn) * a.g_norm;
    double *Aw = a.A + (size_t)(3 * F + 4) * (3 * F + 4);                        // the factorisation works on a copy: A itself lives on into the next pass
    for (int k = 0; k < 4; k++) {
        const double g0n = sqrt(dot(g0, g0));
        const V3 av = g0 * (1.0 / g0n);
        V3 tmp = v3(0, 0, 1);
        if (av.x == tmp.x && av.y == tmp.y && av.z == tmp.z) tmp = v3(1, 0, 0);
        V3 bv = tmp - av * dot(av, tmp);
        bv = bv * (1.0 / sqrt(dot(bv, bv)));
        const V3 cv = cross(av, bv);
        VIWB_SYNC();
        if (tid == 0)
            for (int i = 0; i + 1 < F; i++) { align_rows(a, i, true, g0, bv, cv, ta, tb, rows, 9); align_accumulate(a.A, a.b, n, i, ta, tb, rows, 9); }
        VIWB_SYNC();
        for (int e = tid; e < n * n; e += nt) { a.A[e] *= 1000.0; Aw[e] = a.A[e]; }
        for (int e = tid; e < n; e += nt) a.b[e] *= 1000.0;
        VIWB_SYNC();
        ldlt_solve_block(Aw, n, a.b, a.x, a.perm, tid, nt);
        const V3 gn1 = g0 + bv * a.x[n - 3] + cv * a.x[n - 2];
        g0 = gn1 * (1.0 / sqrt(dot(gn1, gn1))) * a.g_norm;
    }
    const double s = a.x[n - 1] / 100.0;
    VIWB_SYNC();
    if (tid == 0) { a.out[0] = s >= 0.0 ? 1.0 : 0.0; st3(a.out + 1, g0); a.out[4] = (double)n; }
    for (int e = tid; e < n - 1; e += nt) a.out[5 + e] = a.x[e];
    if (tid == 0) a.out[5 + n - 1] = s;
}

#ifndef VIWB_HOST_EMU
__global__ void gyro_bias_kernel(GyroBiasArgs a) { gyro_bias_block(a, threadIdx.x, blockDim.x); }
__global__ void align_kernel(AlignArgs a) { align_block(a, threadIdx.x, blockDim.x); }
#endif

}  // namespace viwb

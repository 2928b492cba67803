// kernels_lk.cuh -- pyramidal Lucas-Kanade tracker replacing cv::calcOpticalFlowPyrLK at the call sites
// featureTracker/feature_tracker.cpp:125-127,136,139,145-146,240,244, plus the status post-processing of
// FeatureTracker::trackImage (:141-162, :245-251).
//
// The arithmetic restates OpenCV's lkpyramid.cpp / pyramids.cpp (third party, not under /root/reference; oracle =
// cv2 4.13 in this image): 8-bit pyrDown [1 4 6 4 1]^2 with (+128)>>8 rounding and REFLECT_101, un-normalised
// 3x3 Scharr into int16, 14-bit fixed-point bilinear weights, patch intensities kept with 5 fractional bits,
// float Gauss-Newton updates, eps^2 / oscillation stopping rules, status and L1 error semantics.
//
// Device layout: no derivative images and no padded copies are materialised.  One warp tracks one point; per level
// it stages the 24x24 source patch in shared memory (reflect-101 at the image border), derives the 22x22 Scharr
// samples from it (zero outside the image, like OpenCV's constant-border derivative buffer), interpolates the 21x21
// template + gradient into registers (14 samples per lane), and stages a 32x32 region of the search image that the
// Gauss-Newton iterations sample from shared memory (restaged only if the estimate leaves it).  Work is described by
// task tables in HBM (one LkArgs per {stream, direction}), so one launch covers every camera stream of a batch.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <vector>
#include "vmath.cuh"

namespace viwb {

enum { LK_WIN = 21, LK_HALF = 10, LK_MAXLVL = 4, LK_PATCH = 24, LK_DPATCH = 22 };

struct LkImage { const uint8_t *img[LK_MAXLVL]; int w[LK_MAXLVL], h[LK_MAXLVL], stride[LK_MAXLVL]; };
struct LkArgs {
    LkImage I, J;             // template image and search image
    const float *prev_pts; float *next_pts; uint8_t *status; float *err;
    const int *n_dev;         // optional device-side point count (batched streams); else n
    int n, max_level, max_iter, flags; float eps2, min_eig;
};

VIWB_HD int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
VIWB_HD int lk_pix(const LkImage &im, int l, int x, int y) { return im.img[l][(size_t)reflect101(y, im.h[l]) * im.stride[l] + reflect101(x, im.w[l])]; }
VIWB_HD int cv_round_f(float v) { return (int)lrintf(v); }                 // cvRound: round half to even
VIWB_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }   // CV_DESCALE

// ---- pyrDown: dst (dw x dh) from src (sw x sh)
struct PyrArgs { const uint8_t *src; uint8_t *dst; int sw, sh, sstride, dw, dh, dstride; };
VIWB_D void pyr_down_item(const PyrArgs &a, int idx) {
    if (idx >= a.dw * a.dh) return;
    const int x = idx % a.dw, y = idx / a.dw;
    int acc = 0;
    const int wgt[5] = {1, 4, 6, 4, 1};
    if (x >= 1 && 2 * x + 2 < a.sw) {      // interior columns: no reflection on x
        for (int dy = -2; dy <= 2; dy++) {
            const uint8_t *row = a.src + (size_t)reflect101(2 * y + dy, a.sh) * a.sstride + 2 * x;
            acc += wgt[dy + 2] * (row[-2] + row[2] + 4 * (row[-1] + row[1]) + 6 * row[0]);
        }
    } else {
        for (int dy = -2; dy <= 2; dy++) {
            const uint8_t *row = a.src + (size_t)reflect101(2 * y + dy, a.sh) * a.sstride;
            int r = 0;
            for (int dx = -2; dx <= 2; dx++) r += wgt[dx + 2] * row[reflect101(2 * x + dx, a.sw)];
            acc += wgt[dy + 2] * r;
        }
    }
    a.dst[(size_t)y * a.dstride + x] = (uint8_t)((acc + 128) >> 8);
}

// ---- the tracker: one warp per point, no block-level barriers.
// Per level the warp stages the 24x24 source patch and its 22x22 Scharr samples in its own shared-memory slice, then every
// lane keeps its 14 template / gradient samples in registers; a Gauss-Newton iteration is 14 bilinear samples of the search
// image per lane (straight from L1), integer products, and one exact 64-bit warp reduction.
#ifdef VIWB_HOST_EMU
enum { LK_W = 1 };
#else
enum { LK_W = 32 };
#endif
#ifndef LK_MINB
#define LK_MINB 6
#endif
enum { LK_EPL = (441 + LK_W - 1) / LK_W, LK_PPB = 4, LK_WARP_SMEM = (576 + 968) * 2 + 32 * 36 + 16 };     // elements per lane, points per block, bytes per warp
VIWB_HD size_t lk_smem_bytes(int warps) { return (size_t)warps * LK_WARP_SMEM; }

VIWB_D void lk_warp_sum3(long long &a, long long &b, long long &c) {
#ifndef VIWB_HOST_EMU
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
#else
    (void)a; (void)b; (void)c;
#endif
}
VIWB_D void lk_warp_sum2(long long &a, long long &b) {
#ifndef VIWB_HOST_EMU
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
#else
    (void)a; (void)b;
#endif
}
VIWB_HD bool lk_outside(int ix, int iy, int cols, int rows) { return ix < -LK_WIN || ix >= cols || iy < -LK_WIN || iy >= rows; }

// the search-image region a warp keeps in shared memory: (22 + 2*LK_SLACK)^2 bytes around the current estimate, reflect-101
// resolved while staging, so the Gauss-Newton loop has a single branch-free sampling path
enum { LK_SLACK = 5, LK_JR = 22 + 2 * LK_SLACK, LK_JS = LK_JR + 4 };
VIWB_D void lk_stage_search(uint8_t *jbuf, const uint8_t *jimg, int jstr, int jc, int jr, int jx0, int jy0, int lane) {
    VIWB_SYNCWARP();
    const bool in = jx0 >= 0 && jy0 >= 0 && jx0 + LK_JR <= jc && jy0 + LK_JR <= jr;
    for (int e = lane; e < LK_JR * LK_JR; e += LK_W) {
        const int y = e / LK_JR, x = e - y * LK_JR;
        jbuf[y * LK_JS + x] = in ? jimg[(size_t)(jy0 + y) * jstr + jx0 + x] : jimg[(size_t)reflect101(jy0 + y, jr) * jstr + reflect101(jx0 + x, jc)];
    }
    VIWB_SYNCWARP();
}

VIWB_D void lk_track_warp(const LkArgs &a, int pt, int lane, unsigned char *smem_raw) {
    short *patch = (short *)smem_raw, *dpatch = patch + 576;
    uint8_t *jbuf = (uint8_t *)(dpatch + 968);
    const int npts = a.n_dev ? *a.n_dev : a.n;
    if (pt >= npts) return;
    const int max_level = a.max_level, max_iter = a.max_iter, flags = a.flags;
    const float eps2 = a.eps2, min_eig = a.min_eig;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float px0 = a.prev_pts[2 * pt], py0 = a.prev_pts[2 * pt + 1];
    bool status = true; float errv = 0.f;
    float npx = 0.f, npy = 0.f;     // nextPts[ptidx] (window centre coordinates)
    int Iv[LK_EPL], dxy[LK_EPL];    // template intensity (5 fractional bits) and packed (dx | dy << 16) per owned element
    int eoff[LK_EPL];               // offset of the element inside the staged search region
#pragma unroll
    for (int k = 0; k < LK_EPL; k++) { const int e = lane + LK_W * k, y = e / 21, x = e - y * 21; eoff[k] = y * LK_JS + x; }
    for (int level = max_level; level >= 0; level--) {
        const float sc = (float)(1. / (1 << level));
        float ppx = px0 * sc, ppy = py0 * sc;
        float nx, ny;
        if (level == max_level) {
            if (flags & 4) { nx = a.next_pts[2 * pt] * sc; ny = a.next_pts[2 * pt + 1] * sc; } else { nx = ppx; ny = ppy; }
        } else { nx = npx * 2.f; ny = npy * 2.f; }
        npx = nx; npy = ny;
        ppx -= (float)LK_HALF; ppy -= (float)LK_HALF;
        const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
        const int cols = a.I.w[level], rows = a.I.h[level];
        if (lk_outside(ipx, ipy, cols, rows)) { if (level == 0) { status = false; errv = 0.f; } continue; }
        // stage the source patch [ipy-1, ipy+22] x [ipx-1, ipx+22]
        VIWB_SYNCWARP();
        {
            const uint8_t *img = a.I.img[level]; const int str = a.I.stride[level];
            const bool in = ipx >= 1 && ipy >= 1 && ipx + 23 <= cols && ipy + 23 <= rows;
            for (int e = lane; e < LK_PATCH * LK_PATCH; e += LK_W) {
                const int y = e / LK_PATCH, x = e - y * LK_PATCH;
                patch[e] = in ? (short)img[(size_t)(ipy - 1 + y) * str + ipx - 1 + x]
                              : (short)img[(size_t)reflect101(ipy - 1 + y, rows) * str + reflect101(ipx - 1 + x, cols)];
            }
        }
        VIWB_SYNCWARP();
        // Scharr samples at [ipy, ipy+21] x [ipx, ipx+21]; zero outside the image
        for (int e = lane; e < LK_DPATCH * LK_DPATCH; e += LK_W) {
            const int y = e / LK_DPATCH, x = e - y * LK_DPATCH, gx = ipx + x, gy = ipy + y;
            int ddx = 0, ddy = 0;
            if (gx >= 0 && gy >= 0 && gx < cols && gy < rows) {
                const short *p0 = patch + y * LK_PATCH + x, *p1 = p0 + LK_PATCH, *p2 = p1 + LK_PATCH;
                ddx = (p0[2] + p2[2]) * 3 + p1[2] * 10 - ((p0[0] + p2[0]) * 3 + p1[0] * 10);
                ddy = ((p2[2] - p0[2]) + (p2[0] - p0[0])) * 3 + (p2[1] - p0[1]) * 10;
            }
            dpatch[2 * e] = (short)ddx; dpatch[2 * e + 1] = (short)ddy;
        }
        VIWB_SYNCWARP();
        const float fa = ppx - ipx, fb = ppy - ipy;
        const int iw00 = cv_round_f((1.f - fa) * (1.f - fb) * (1 << 14)), iw01 = cv_round_f(fa * (1.f - fb) * (1 << 14));
        const int iw10 = cv_round_f((1.f - fa) * fb * (1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
        long long A11 = 0, A12 = 0, A22 = 0;
#pragma unroll
        for (int k = 0; k < LK_EPL; k++) {
            const int e = lane + LK_W * k;
            if (e < 441) {
                const int y = e / 21, x = e - y * 21;
                const short *p = patch + (y + 1) * LK_PATCH + x + 1, *d = dpatch + 2 * (y * LK_DPATCH + x);
                const int ival = descale(p[0] * iw00 + p[1] * iw01 + p[LK_PATCH] * iw10 + p[LK_PATCH + 1] * iw11, 9);
                const int ix = descale(d[0] * iw00 + d[2] * iw01 + d[2 * LK_DPATCH] * iw10 + d[2 * LK_DPATCH + 2] * iw11, 14);
                const int iy = descale(d[1] * iw00 + d[3] * iw01 + d[2 * LK_DPATCH + 1] * iw10 + d[2 * LK_DPATCH + 3] * iw11, 14);
                Iv[k] = ival; dxy[k] = (ix & 0xffff) | (iy << 16);
                A11 += (long long)(ix * ix); A12 += (long long)(ix * iy); A22 += (long long)(iy * iy);
            } else { Iv[k] = 0; dxy[k] = 0; }
        }
        lk_warp_sum3(A11, A12, A22);
        const float fA11 = (float)A11 * FLT_SCALE, fA12 = (float)A12 * FLT_SCALE, fA22 = (float)A22 * FLT_SCALE;
        float D = fA11 * fA22 - fA12 * fA12;
        const float minEig = (fA22 + fA11 - sqrtf((fA11 - fA22) * (fA11 - fA22) + 4.f * fA12 * fA12)) / (2 * LK_WIN * LK_WIN);
        if (minEig < min_eig || D < 1.1920929e-07f) { if (level == 0) status = false; continue; }
        D = 1.f / D;
        nx -= (float)LK_HALF; ny -= (float)LK_HALF;
        float pdx = 0.f, pdy = 0.f;
        const int jc = a.J.w[level], jr = a.J.h[level], jstr = a.J.stride[level];
        const uint8_t *jimg = a.J.img[level];
        int jx0 = (int)floorf(nx) - LK_SLACK, jy0 = (int)floorf(ny) - LK_SLACK;       // origin of the staged search region
        bool staged = false;
        for (int j = 0; j < max_iter; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (lk_outside(inx, iny, jc, jr)) { if (level == 0) status = false; break; }
            if (!staged || inx < jx0 || iny < jy0 || inx + 22 > jx0 + LK_JR || iny + 22 > jy0 + LK_JR) {
                jx0 = inx - LK_SLACK; jy0 = iny - LK_SLACK;
                lk_stage_search(jbuf, jimg, jstr, jc, jr, jx0, jy0, lane);
                staged = true;
            }
            const float ja = nx - inx, jb = ny - iny;
            const int w00 = cv_round_f((1.f - ja) * (1.f - jb) * (1 << 14)), w01 = cv_round_f(ja * (1.f - jb) * (1 << 14));
            const int w10 = cv_round_f((1.f - ja) * jb * (1 << 14)), w11 = (1 << 14) - w00 - w01 - w10;
            const uint8_t *q0 = jbuf + (iny - jy0) * LK_JS + (inx - jx0);
            long long b1 = 0, b2 = 0;
#pragma unroll
            for (int k = 0; k < LK_EPL; k++) {
                if (lane + LK_W * k < 441) {
                    const uint8_t *q = q0 + eoff[k];
                    const int diff = descale(q[0] * w00 + q[1] * w01 + q[LK_JS] * w10 + q[LK_JS + 1] * w11, 9) - Iv[k];
                    b1 += (long long)(diff * (int)(short)(dxy[k] & 0xffff)); b2 += (long long)(diff * (dxy[k] >> 16));
                }
            }
            lk_warp_sum2(b1, b2);
            const float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
            const float dx = (fA12 * fb2 - fA22 * fb1) * D, dy = (fA12 * fb1 - fA11 * fb2) * D;
            nx += dx; ny += dy;
            npx = nx + (float)LK_HALF; npy = ny + (float)LK_HALF;
            if (dx * dx + dy * dy <= eps2) break;
            if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) { npx -= dx * 0.5f; npy -= dy * 0.5f; break; }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {   // L1 error of the final patch (flags without OPTFLOW_LK_GET_MIN_EIGENVALS)
            const float ex = npx - (float)LK_HALF, ey = npy - (float)LK_HALF;
            const int inx = (int)floorf(ex), iny = (int)floorf(ey);
            if (lk_outside(inx, iny, jc, jr)) { status = false; }
            else {
                if (!staged || inx < jx0 || iny < jy0 || inx + 22 > jx0 + LK_JR || iny + 22 > jy0 + LK_JR) {
                    jx0 = inx - LK_SLACK; jy0 = iny - LK_SLACK;
                    lk_stage_search(jbuf, jimg, jstr, jc, jr, jx0, jy0, lane);
                }
                const float ja = ex - inx, jb = ey - iny;
                const int w00 = cv_round_f((1.f - ja) * (1.f - jb) * (1 << 14)), w01 = cv_round_f(ja * (1.f - jb) * (1 << 14));
                const int w10 = cv_round_f((1.f - ja) * jb * (1 << 14)), w11 = (1 << 14) - w00 - w01 - w10;
                const uint8_t *q0 = jbuf + (iny - jy0) * LK_JS + (inx - jx0);
                long long ev = 0, d1 = 0;
#pragma unroll
                for (int k = 0; k < LK_EPL; k++) {
                    if (lane + LK_W * k < 441) {
                        const uint8_t *q = q0 + eoff[k];
                        const int diff = descale(q[0] * w00 + q[1] * w01 + q[LK_JS] * w10 + q[LK_JS + 1] * w11, 9) - Iv[k];
                        ev += (long long)(diff < 0 ? -diff : diff);
                    }
                }
                lk_warp_sum2(ev, d1);
                errv = (float)ev * (1.f / (32 * LK_WIN * LK_WIN));
            }
        }
    }
    if (lane == 0) { a.next_pts[2 * pt] = npx; a.next_pts[2 * pt + 1] = npy; a.status[pt] = status ? 1 : 0; if (a.err) a.err[pt] = errv; }
}

// ---- status post-processing of trackImage: round trip <= 0.5 px and the 1-px border test after cvRound
struct PostArgs { const float *pts_a, *pts_b, *pts_back; uint8_t *status; const uint8_t *status_back; const int *n_dev; int n, w, h, mode, flow_back; };
VIWB_HD bool lk_in_border(float x, float y, int w, int h) {
    const int ix = cv_round_f(x), iy = cv_round_f(y);
    return 1 <= ix && ix < w - 1 && 1 <= iy && iy < h - 1;
}
VIWB_D void lk_post_item(const PostArgs &a, int i) {
    const int n = a.n_dev ? *a.n_dev : a.n;
    if (i >= n) return;
    int st = a.status[i];
    if (a.flow_back) {
        const double dx = (double)(a.pts_a[2 * i] - a.pts_back[2 * i]), dy = (double)(a.pts_a[2 * i + 1] - a.pts_back[2 * i + 1]);
        bool ok = st && a.status_back[i] && sqrt(dx * dx + dy * dy) <= 0.5;
        if (a.mode == 1) ok = ok && lk_in_border(a.pts_b[2 * i], a.pts_b[2 * i + 1], a.w, a.h);
        st = ok ? 1 : 0;
    }
    if (a.mode == 0 && st && !lk_in_border(a.pts_b[2 * i], a.pts_b[2 * i + 1], a.w, a.h)) st = 0;
    a.status[i] = (uint8_t)st;
}

#ifndef VIWB_HOST_EMU
// task-table kernels: blockIdx.y selects the task
__global__ void pyr_down_tasks_kernel(const PyrArgs *t) { pyr_down_item(t[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void lk_post_tasks_kernel(const PostArgs *t) { lk_post_item(t[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(32 * LK_PPB, LK_MINB) lk_track_tasks_kernel(const LkArgs *t) {
    extern __shared__ unsigned char lk_smem[];
    const int warp = threadIdx.x >> 5;
    lk_track_warp(t[blockIdx.y], blockIdx.x * LK_PPB + warp, threadIdx.x & 31, lk_smem + (size_t)warp * LK_WARP_SMEM);
}
#endif

}  // namespace viwb

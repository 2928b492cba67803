// kernels_lk.cuh -- pyramidal Lucas-Kanade tracker replacing cv::calcOpticalFlowPyrLK at the call sites
// featureTracker/feature_tracker.cpp:125-127,136,139,145-146,240,244, plus the status post-processing of
// FeatureTracker::trackImage (:141-162, :245-251).
//
// The arithmetic restates OpenCV's lkpyramid.cpp / pyramids.cpp (third party, not under /root/reference; oracle =
// cv2 4.13 in this image): 8-bit pyrDown [1 4 6 4 1]^2 with (+128)>>8 rounding and REFLECT_101, un-normalised
// 3x3 Scharr into int16, 14-bit fixed-point bilinear weights, patch intensities kept with 5 fractional bits,
// float Gauss-Newton updates, eps^2 / oscillation stopping rules, status and L1 error semantics.  Borders are
// resolved on the fly (reflect-101 for intensities, zero for derivatives) instead of materialising padded copies.
// Launch shape: one block per tracked point; the 21x21 template lives in shared memory for all iterations.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <vector>
#include "vmath.cuh"

namespace viwb {

enum { LK_WIN = 21, LK_HALF = 10, LK_MAXLVL = 4, LK_NT = 64 };

struct LkImage { const uint8_t *img[LK_MAXLVL]; const short *deriv[LK_MAXLVL]; int w[LK_MAXLVL], h[LK_MAXLVL], stride[LK_MAXLVL]; };
struct LkArgs {
    LkImage I, J;             // template image (with derivatives) and search image
    const float *prev_pts; float *next_pts; uint8_t *status; float *err;
    int n, max_level, max_iter, flags; float eps2, min_eig;
};

VIWB_HD int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
VIWB_HD int lk_pix(const LkImage &im, int l, int x, int y) { return im.img[l][(size_t)reflect101(y, im.h[l]) * im.stride[l] + reflect101(x, im.w[l])]; }
VIWB_HD int lk_der(const LkImage &im, int l, int x, int y, int c) {
    if (x < 0 || y < 0 || x >= im.w[l] || y >= im.h[l]) return 0;
    return im.deriv[l][((size_t)y * im.w[l] + x) * 2 + c];
}
VIWB_HD int cv_round_f(float v) { return (int)lrintf(v); }                 // cvRound: round half to even
VIWB_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }   // CV_DESCALE

// ---- pyrDown: dst (dw x dh) from src (sw x sh)
struct PyrArgs { const uint8_t *src; uint8_t *dst; int sw, sh, sstride, dw, dh, dstride; };
VIWB_D void pyr_down_item(const PyrArgs &a, int idx) {
    if (idx >= a.dw * a.dh) return;
    const int x = idx % a.dw, y = idx / a.dw;
    int acc = 0;
    const int wgt[5] = {1, 4, 6, 4, 1};
    for (int dy = -2; dy <= 2; dy++) {
        const uint8_t *row = a.src + (size_t)reflect101(2 * y + dy, a.sh) * a.sstride;
        int r = 0;
        for (int dx = -2; dx <= 2; dx++) r += wgt[dx + 2] * row[reflect101(2 * x + dx, a.sw)];
        acc += wgt[dy + 2] * r;
    }
    a.dst[(size_t)y * a.dstride + x] = (uint8_t)((acc + 128) >> 8);
}
// ---- Scharr derivative (calcSharrDeriv): int16 interleaved (dx, dy)
struct ScharrArgs { const uint8_t *src; short *dst; int w, h, stride; };
VIWB_D void scharr_item(const ScharrArgs &a, int idx) {
    if (idx >= a.w * a.h) return;
    const int x = idx % a.w, y = idx / a.w;
    const uint8_t *r0 = a.src + (size_t)reflect101(y - 1, a.h) * a.stride, *r1 = a.src + (size_t)y * a.stride, *r2 = a.src + (size_t)reflect101(y + 1, a.h) * a.stride;
    const int xm = reflect101(x - 1, a.w), xp = reflect101(x + 1, a.w);
    const int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10, t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
    const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
    a.dst[(size_t)idx * 2] = (short)(t0p - t0m);
    a.dst[(size_t)idx * 2 + 1] = (short)((t1p + t1m) * 3 + t1c * 10);
}

// ---- the tracker: one block per point.  smem: short Iptr[441], short dI[882], double red[3*nt], float bc[8]
VIWB_HD size_t lk_smem_bytes(int nt) { return 441 * 2 + 882 * 2 + 6 + (size_t)3 * nt * 8 + 64; }

VIWB_D void lk_sum3(double &a, double &b, double &c, int tid, int nt, double *red) {
    red[tid] = a; red[nt + tid] = b; red[2 * nt + tid] = c;
    VIWB_SYNC();
    for (int s = nt >> 1; s > 0; s >>= 1) { if (tid < s) { red[tid] += red[tid + s]; red[nt + tid] += red[nt + tid + s]; red[2 * nt + tid] += red[2 * nt + tid + s]; } VIWB_SYNC(); }
    a = red[0]; b = red[nt]; c = red[2 * nt];
    VIWB_SYNC();
}
VIWB_HD bool lk_outside(int ix, int iy, int cols, int rows) { return ix < -LK_WIN || ix >= cols || iy < -LK_WIN || iy >= rows; }

VIWB_D void lk_track_block(const LkArgs &a, int pt, int tid, int nt, unsigned char *smem_raw) {
    short *Iptr = (short *)smem_raw, *dI = Iptr + 441;
    double *red = (double *)(smem_raw + ((441 * 2 + 882 * 2 + 7) / 8) * 8);
    if (pt >= a.n) return;
    const float FLT_SCALE = 1.f / (1 << 20);
    bool status = true; float errv = 0.f;
    float npx = 0.f, npy = 0.f;     // nextPts[ptidx] (window centre coordinates)
    for (int level = a.max_level; level >= 0; level--) {
        const float sc = (float)(1. / (1 << level));
        float ppx = a.prev_pts[2 * pt] * sc, ppy = a.prev_pts[2 * pt + 1] * sc;
        float nx, ny;
        if (level == a.max_level) {
            if (a.flags & 4) { nx = a.next_pts[2 * pt] * sc; ny = a.next_pts[2 * pt + 1] * sc; } else { nx = ppx; ny = ppy; }
        } else { nx = npx * 2.f; ny = npy * 2.f; }
        npx = nx; npy = ny;
        ppx -= (float)LK_HALF; ppy -= (float)LK_HALF;
        const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
        const int cols = a.I.w[level], rows = a.I.h[level];
        if (lk_outside(ipx, ipy, cols, rows)) { if (level == 0) { status = false; errv = 0.f; } continue; }
        {
            const float fa = ppx - ipx, fb = ppy - ipy;
            const int iw00 = cv_round_f((1.f - fa) * (1.f - fb) * (1 << 14)), iw01 = cv_round_f(fa * (1.f - fb) * (1 << 14));
            const int iw10 = cv_round_f((1.f - fa) * fb * (1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
            double A11 = 0, A12 = 0, A22 = 0;
            for (int e = tid; e < 441; e += nt) {
                const int y = e / 21, x = e % 21, sx = ipx + x, sy = ipy + y;
                const int ival = descale(lk_pix(a.I, level, sx, sy) * iw00 + lk_pix(a.I, level, sx + 1, sy) * iw01 + lk_pix(a.I, level, sx, sy + 1) * iw10 + lk_pix(a.I, level, sx + 1, sy + 1) * iw11, 9);
                const int ix = descale(lk_der(a.I, level, sx, sy, 0) * iw00 + lk_der(a.I, level, sx + 1, sy, 0) * iw01 + lk_der(a.I, level, sx, sy + 1, 0) * iw10 + lk_der(a.I, level, sx + 1, sy + 1, 0) * iw11, 14);
                const int iy = descale(lk_der(a.I, level, sx, sy, 1) * iw00 + lk_der(a.I, level, sx + 1, sy, 1) * iw01 + lk_der(a.I, level, sx, sy + 1, 1) * iw10 + lk_der(a.I, level, sx + 1, sy + 1, 1) * iw11, 14);
                Iptr[e] = (short)ival; dI[2 * e] = (short)ix; dI[2 * e + 1] = (short)iy;
                A11 += (double)(ix * ix); A12 += (double)(ix * iy); A22 += (double)(iy * iy);
            }
            lk_sum3(A11, A12, A22, tid, nt, red);
            const float fA11 = (float)A11 * FLT_SCALE, fA12 = (float)A12 * FLT_SCALE, fA22 = (float)A22 * FLT_SCALE;
            float D = fA11 * fA22 - fA12 * fA12;
            const float minEig = (fA22 + fA11 - sqrtf((fA11 - fA22) * (fA11 - fA22) + 4.f * fA12 * fA12)) / (2 * LK_WIN * LK_WIN);
            if (minEig < a.min_eig || D < 1.1920929e-07f) { if (level == 0) status = false; continue; }
            D = 1.f / D;
            nx -= (float)LK_HALF; ny -= (float)LK_HALF;
            float pdx = 0.f, pdy = 0.f;
            const int jc = a.J.w[level], jr = a.J.h[level];
            for (int j = 0; j < a.max_iter; j++) {
                const int inx = (int)floorf(nx), iny = (int)floorf(ny);
                if (lk_outside(inx, iny, jc, jr)) { if (level == 0) status = false; break; }
                const float ja = nx - inx, jb = ny - iny;
                const int w00 = cv_round_f((1.f - ja) * (1.f - jb) * (1 << 14)), w01 = cv_round_f(ja * (1.f - jb) * (1 << 14));
                const int w10 = cv_round_f((1.f - ja) * jb * (1 << 14)), w11 = (1 << 14) - w00 - w01 - w10;
                double b1 = 0, b2 = 0, dummy = 0;
                for (int e = tid; e < 441; e += nt) {
                    const int y = e / 21, x = e % 21, sx = inx + x, sy = iny + y;
                    const int diff = descale(lk_pix(a.J, level, sx, sy) * w00 + lk_pix(a.J, level, sx + 1, sy) * w01 + lk_pix(a.J, level, sx, sy + 1) * w10 + lk_pix(a.J, level, sx + 1, sy + 1) * w11, 9) - Iptr[e];
                    b1 += (double)(diff * dI[2 * e]); b2 += (double)(diff * dI[2 * e + 1]);
                }
                lk_sum3(b1, b2, dummy, tid, nt, red);
                const float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
                const float dx = (fA12 * fb2 - fA22 * fb1) * D, dy = (fA12 * fb1 - fA11 * fb2) * D;
                nx += dx; ny += dy;
                npx = nx + (float)LK_HALF; npy = ny + (float)LK_HALF;
                if (dx * dx + dy * dy <= a.eps2) break;
                if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) { npx -= dx * 0.5f; npy -= dy * 0.5f; break; }
                pdx = dx; pdy = dy;
            }
            if (status && level == 0) {   // L1 error of the final patch (flags without OPTFLOW_LK_GET_MIN_EIGENVALS)
                const float ex = npx - (float)LK_HALF, ey = npy - (float)LK_HALF;
                const int inx = (int)floorf(ex), iny = (int)floorf(ey);
                if (lk_outside(inx, iny, jc, jr)) { status = false; }
                else {
                    const float ja = ex - inx, jb = ey - iny;
                    const int w00 = cv_round_f((1.f - ja) * (1.f - jb) * (1 << 14)), w01 = cv_round_f(ja * (1.f - jb) * (1 << 14));
                    const int w10 = cv_round_f((1.f - ja) * jb * (1 << 14)), w11 = (1 << 14) - w00 - w01 - w10;
                    double ev = 0, d1 = 0, d2 = 0;
                    for (int e = tid; e < 441; e += nt) {
                        const int y = e / 21, x = e % 21, sx = inx + x, sy = iny + y;
                        const int diff = descale(lk_pix(a.J, level, sx, sy) * w00 + lk_pix(a.J, level, sx + 1, sy) * w01 + lk_pix(a.J, level, sx, sy + 1) * w10 + lk_pix(a.J, level, sx + 1, sy + 1) * w11, 9) - Iptr[e];
                        ev += (double)(diff < 0 ? -diff : diff);
                    }
                    lk_sum3(ev, d1, d2, tid, nt, red);
                    errv = (float)ev * (1.f / (32 * LK_WIN * LK_WIN));
                }
            }
        }
        VIWB_SYNC();
    }
    if (tid == 0) { a.next_pts[2 * pt] = npx; a.next_pts[2 * pt + 1] = npy; a.status[pt] = status ? 1 : 0; if (a.err) a.err[pt] = errv; }
}

// ---- status post-processing of trackImage: round trip <= 0.5 px and the 1-px border test after cvRound
struct PostArgs { const float *pts_a, *pts_b, *pts_back; uint8_t *status; const uint8_t *status_back; int n, w, h, mode, flow_back; };
VIWB_HD bool lk_in_border(float x, float y, int w, int h) {
    const int ix = cv_round_f(x), iy = cv_round_f(y);
    return 1 <= ix && ix < w - 1 && 1 <= iy && iy < h - 1;
}
VIWB_D void lk_post_item(const PostArgs &a, int i) {
    if (i >= a.n) return;
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

// ====================================================================================== host side
#ifdef VIWB_HOST_EMU
#define LK_LAUNCH_ITEMS(fn, args, items, stream) do { for (int i_ = 0; i_ < (items); i_++) fn(args, i_); } while (0)
static void lk_launch_track(const LkArgs &a, void *) { std::vector<unsigned char> sm(lk_smem_bytes(1) + 64); for (int p = 0; p < a.n; p++) lk_track_block(a, p, 0, 1, sm.data()); }
static int lk_malloc(void **p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xFF, n ? n : 1); return *p ? 0 : 1; }
static void lk_free(void *p) { free(p); }
static int lk_h2d(void *d, const void *h, size_t n, void *) { memcpy(d, h, n); return 0; }
static int lk_d2h(void *h, const void *d, size_t n, void *) { memcpy(h, d, n); return 0; }
static int lk_h2d_2d(void *d, size_t dp, const void *h, size_t hp, size_t w, size_t hh, void *) { for (size_t y = 0; y < hh; y++) memcpy((char *)d + y * dp, (const char *)h + y * hp, w); return 0; }
static int lk_sync(void *) { return 0; }
typedef void *lk_stream_t;
#else
__global__ void pyr_down_kernel(PyrArgs a) { pyr_down_item(a, blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void scharr_kernel(ScharrArgs a) { scharr_item(a, blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void lk_post_kernel(PostArgs a) { lk_post_item(a, blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(LK_NT) lk_track_kernel(LkArgs a) { extern __shared__ unsigned char lk_smem[]; lk_track_block(a, blockIdx.x, threadIdx.x, blockDim.x, lk_smem); }
#define LK_LAUNCH_ITEMS(fn, args, items, stream) do { if ((items) > 0) fn##_k_sel(args, items, stream); } while (0)
static void pyr_down_item_k_sel(const PyrArgs &a, int items, cudaStream_t s) { pyr_down_kernel<<<(items + 255) / 256, 256, 0, s>>>(a); }
static void scharr_item_k_sel(const ScharrArgs &a, int items, cudaStream_t s) { scharr_kernel<<<(items + 255) / 256, 256, 0, s>>>(a); }
static void lk_post_item_k_sel(const PostArgs &a, int items, cudaStream_t s) { lk_post_kernel<<<(items + 127) / 128, 128, 0, s>>>(a); }
static void lk_launch_track(const LkArgs &a, cudaStream_t s) { if (a.n > 0) lk_track_kernel<<<a.n, LK_NT, lk_smem_bytes(LK_NT), s>>>(a); }
static int lk_malloc(void **p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
static void lk_free(void *p) { cudaFree(p); }
static int lk_h2d(void *d, const void *h, size_t n, cudaStream_t s) { return (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s); }
static int lk_d2h(void *h, const void *d, size_t n, cudaStream_t s) { return (int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s); }
static int lk_h2d_2d(void *d, size_t dp, const void *h, size_t hp, size_t w, size_t hh, cudaStream_t s) { return (int)cudaMemcpy2DAsync(d, dp, h, hp, w, hh, cudaMemcpyHostToDevice, s); }
static int lk_sync(cudaStream_t s) { return (int)cudaStreamSynchronize(s); }
typedef cudaStream_t lk_stream_t;
#endif

// device buffers of one image pyramid (+ derivatives), cached across calls
struct LkPyramid {
    uint8_t *img[LK_MAXLVL]; short *deriv[LK_MAXLVL]; int w[LK_MAXLVL], h[LK_MAXLVL], stride[LK_MAXLVL]; int cap_w, cap_h, levels;
    LkPyramid() { memset(this, 0, sizeof *this); }
};
struct LkPool { LkPyramid a, b; float *pts[3]; uint8_t *st[2]; float *err; int cap_n; LkPool() : cap_n(0) { pts[0] = pts[1] = pts[2] = nullptr; st[0] = st[1] = nullptr; err = nullptr; } };
static LkPool g_lk_pool;

static int lk_pyr_alloc(LkPyramid &p, int w, int h) {
    if (p.cap_w == w && p.cap_h == h) return 0;
    for (int l = 0; l < LK_MAXLVL; l++) { if (p.img[l]) lk_free(p.img[l]); if (p.deriv[l]) lk_free(p.deriv[l]); p.img[l] = nullptr; p.deriv[l] = nullptr; }
    int cw = w, ch = h;
    for (int l = 0; l < LK_MAXLVL; l++) {
        p.w[l] = cw; p.h[l] = ch; p.stride[l] = (cw + 15) & ~15;
        if (lk_malloc((void **)&p.img[l], (size_t)p.stride[l] * ch)) return 1;
        if (lk_malloc((void **)&p.deriv[l], (size_t)cw * ch * 2 * sizeof(short))) return 1;
        cw = (cw + 1) / 2; ch = (ch + 1) / 2;
    }
    p.cap_w = w; p.cap_h = h;
    return 0;
}
static int lk_pool_pts(int n) {
    LkPool &g = g_lk_pool;
    if (n <= g.cap_n) return 0;
    for (int i = 0; i < 3; i++) { if (g.pts[i]) lk_free(g.pts[i]); if (lk_malloc((void **)&g.pts[i], (size_t)n * 2 * sizeof(float))) return 1; }
    for (int i = 0; i < 2; i++) { if (g.st[i]) lk_free(g.st[i]); if (lk_malloc((void **)&g.st[i], (size_t)n)) return 1; }
    if (g.err) lk_free(g.err); if (lk_malloc((void **)&g.err, (size_t)n * sizeof(float))) return 1;
    g.cap_n = n;
    return 0;
}
static void lk_release(int) {
    LkPool &g = g_lk_pool;
    LkPyramid *ps[2] = {&g.a, &g.b};
    for (LkPyramid *p : ps) { for (int l = 0; l < LK_MAXLVL; l++) { if (p->img[l]) lk_free(p->img[l]); if (p->deriv[l]) lk_free(p->deriv[l]); } memset(p, 0, sizeof *p); }
    for (int i = 0; i < 3; i++) { if (g.pts[i]) lk_free(g.pts[i]); g.pts[i] = nullptr; }
    for (int i = 0; i < 2; i++) { if (g.st[i]) lk_free(g.st[i]); g.st[i] = nullptr; }
    if (g.err) lk_free(g.err); g.err = nullptr; g.cap_n = 0;
}
// number of usable levels: buildOpticalFlowPyramid stops once a level is not larger than the window
static int lk_levels(int w, int h, int max_level) {
    int lv = 0, cw = w, ch = h;
    for (int l = 1; l <= max_level; l++) { cw = (cw + 1) / 2; ch = (ch + 1) / 2; if (cw <= LK_WIN || ch <= LK_WIN) break; lv = l; }
    return lv;
}
// upload level 0, build the pyramid and (optionally) the Scharr derivatives
static int lk_build(LkPyramid &p, const uint8_t *host, int w, int h, int stride, int levels, bool deriv, lk_stream_t s, long long *nl) {
    if (lk_pyr_alloc(p, w, h)) return 1;
    if (lk_h2d_2d(p.img[0], p.stride[0], host, stride, w, h, s)) return 1;
    for (int l = 1; l <= levels; l++) {
        PyrArgs a; a.src = p.img[l - 1]; a.dst = p.img[l]; a.sw = p.w[l - 1]; a.sh = p.h[l - 1]; a.sstride = p.stride[l - 1]; a.dw = p.w[l]; a.dh = p.h[l]; a.dstride = p.stride[l];
        LK_LAUNCH_ITEMS(pyr_down_item, a, a.dw * a.dh, s); (*nl)++;
    }
    if (deriv) for (int l = 0; l <= levels; l++) {
        ScharrArgs a; a.src = p.img[l]; a.dst = p.deriv[l]; a.w = p.w[l]; a.h = p.h[l]; a.stride = p.stride[l];
        LK_LAUNCH_ITEMS(scharr_item, a, a.w * a.h, s); (*nl)++;
    }
    p.levels = levels;
    return 0;
}
static void lk_fill(LkImage &im, const LkPyramid &p) { for (int l = 0; l < LK_MAXLVL; l++) { im.img[l] = p.img[l]; im.deriv[l] = p.deriv[l]; im.w[l] = p.w[l]; im.h[l] = p.h[l]; im.stride[l] = p.stride[l]; } }
static void lk_criteria(int max_iter, float eps, int &mi, float &e2) {
    mi = max_iter < 0 ? 0 : (max_iter > 100 ? 100 : max_iter);
    float e = eps < 0.f ? 0.f : (eps > 10.f ? 10.f : eps);
    e2 = e * e;
}

static int lk_track_host(int device, lk_stream_t s, const uint8_t *prev, const uint8_t *next, int w, int h, int stride, const float *prev_pts,
                         float *next_pts, int n, int max_level, int max_iter, float eps, int flags, float min_eig, uint8_t *status, float *err, long long *nl) {
    (void)device;
    if (n == 0) return 0;
    LkPool &g = g_lk_pool;
    const int levels = lk_levels(w, h, max_level);
    if (lk_build(g.a, prev, w, h, stride, levels, true, s, nl)) return 1;
    if (lk_build(g.b, next, w, h, stride, levels, false, s, nl)) return 1;
    if (lk_pool_pts(n)) return 1;
    if (lk_h2d(g.pts[0], prev_pts, (size_t)n * 8, s)) return 1;
    if (lk_h2d(g.pts[1], (flags & 4) ? next_pts : prev_pts, (size_t)n * 8, s)) return 1;
    LkArgs a; lk_fill(a.I, g.a); lk_fill(a.J, g.b);
    a.prev_pts = g.pts[0]; a.next_pts = g.pts[1]; a.status = g.st[0]; a.err = g.err; a.n = n; a.max_level = levels; a.flags = flags; a.min_eig = min_eig;
    lk_criteria(max_iter, eps, a.max_iter, a.eps2);
    lk_launch_track(a, s); (*nl)++;
    if (lk_d2h(next_pts, g.pts[1], (size_t)n * 8, s)) return 1;
    if (lk_d2h(status, g.st[0], (size_t)n, s)) return 1;
    if (err && lk_d2h(err, g.err, (size_t)n * 4, s)) return 1;
    return lk_sync(s);
}

// forward + (optional) reverse LK sharing the two pyramids, then the reference's status rules, all on the device
static int lk_track_checked_host(int device, lk_stream_t s, const uint8_t *img_a, const uint8_t *img_b, int w, int h, int stride, const float *pts_a,
                                 float *pts_b, int n, int mode, int flow_back, uint8_t *status, long long *nl) {
    (void)device;
    if (n == 0) return 0;
    LkPool &g = g_lk_pool;
    const int levels = lk_levels(w, h, 3);
    if (lk_build(g.a, img_a, w, h, stride, levels, true, s, nl)) return 1;
    if (lk_build(g.b, img_b, w, h, stride, levels, flow_back != 0, s, nl)) return 1;
    if (lk_pool_pts(n)) return 1;
    if (lk_h2d(g.pts[0], pts_a, (size_t)n * 8, s)) return 1;
    if (lk_h2d(g.pts[1], pts_a, (size_t)n * 8, s)) return 1;
    LkArgs a; lk_fill(a.I, g.a); lk_fill(a.J, g.b);
    a.prev_pts = g.pts[0]; a.next_pts = g.pts[1]; a.status = g.st[0]; a.err = g.err; a.n = n; a.max_level = levels; a.flags = 0; a.min_eig = 1e-4f;
    lk_criteria(30, 0.01f, a.max_iter, a.eps2);
    lk_launch_track(a, s); (*nl)++;
    if (flow_back) {
        LkArgs r; lk_fill(r.I, g.b); lk_fill(r.J, g.a);
        r.prev_pts = g.pts[1]; r.next_pts = g.pts[2]; r.status = g.st[1]; r.err = g.err; r.n = n; r.min_eig = 1e-4f;
        lk_criteria(30, 0.01f, r.max_iter, r.eps2);
        if (mode == 0) {   // temporal: maxLevel 1, OPTFLOW_USE_INITIAL_FLOW seeded with prev_pts (feature_tracker.cpp:144-146)
            if (lk_h2d(g.pts[2], pts_a, (size_t)n * 8, s)) return 1;
            r.max_level = levels < 1 ? levels : 1; r.flags = 4;
        } else { r.max_level = levels; r.flags = 0; }            // stereo: maxLevel 3, no initial flow (:244)
        lk_launch_track(r, s); (*nl)++;
    }
    PostArgs p; p.pts_a = g.pts[0]; p.pts_b = g.pts[1]; p.pts_back = g.pts[2]; p.status = g.st[0]; p.status_back = g.st[1]; p.n = n; p.w = w; p.h = h; p.mode = mode; p.flow_back = flow_back;
    LK_LAUNCH_ITEMS(lk_post_item, p, n, s); (*nl)++;
    if (lk_d2h(pts_b, g.pts[1], (size_t)n * 8, s)) return 1;
    if (lk_d2h(status, g.st[0], (size_t)n, s)) return 1;
    return lk_sync(s);
}

}  // namespace viwb

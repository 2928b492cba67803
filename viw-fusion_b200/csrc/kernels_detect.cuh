// kernels_detect.cuh -- the feature-detection half of FeatureTracker::trackImage():
//   FeatureTracker::setMask()                                   featureTracker/feature_tracker.cpp:59-89
//   cv::goodFeaturesToTrack(cur_img, n_pts, MAX_CNT - n, 0.01, MIN_DIST, mask)              feature_tracker.cpp:192
//
// The arithmetic restates OpenCV's published algorithm (third party, not under /root/reference; checker = cv2 4.13 scalar code
// with its SIMD paths off): cornerMinEigenVal = FP32 3x3 Sobel with the 1/(255*12) scale folded into the smoothing
// kernel, REFLECT_101 borders, products dx^2 / dx*dy / dy^2 box-summed over 3x3 in FP64 and narrowed to FP32 (the box filter
// reflects the *product* image at the border, not the source), (a+c) - sqrt((a-c)^2 + b^2); then the masked maximum, THRESH_TOZERO
// at max*quality, strict 3x3 local maxima in rows / columns 1..n-2, descending order (ties: larger address first) and the greedy
// minimum-distance selection over a cell grid.  Every FP32 operation uses the round-to-nearest intrinsics so that nvcc cannot contract
// a multiply-add: the results are the same bits as the scalar CPU code.
//
// Device layout (per camera stream): the image stays where the tracker put it; `mask` u8 [h][w], a candidate list of 64-bit keys
// (ordered eigenvalue bits << 32 | pixel offset) and a cell grid with four packed (x | y << 16) slots per cell -- cells are min_dist
// wide, so a cell can never hold more than four accepted corners.  The eigenvalue image itself is never written to HBM.  Four
// launches cover every stream of a batch (blockIdx.y / task index = stream): order -> mask -> corners -> select.
#pragma once
#include <stdint.h>
#include <math.h>
#include "vmath.cuh"

namespace viwb {

enum { DET_SLOTS = 4, DET_SORT_SMEM = 8192, DET_GRID_SMEM = 1024, DET_BAND = 8, DET_MAXPTS = 1024 };

struct DetArgs {                        // one camera stream
    const unsigned char *img[2];        // candidate image locations (the tracker's two alternating left slots); DetRun::img_sel picks one
    int w, h, stride;
    const unsigned char *base_mask;     // optional fisheye mask [h][w] (feature_tracker.cpp:63 starts from a white image when it is absent)
    const float *pts; const int *track_cnt; const int *n_dev;      // tracked points of the stream (setMask input)
    int radius; const short *hw;        // MIN_DIST and the half widths of the filled circle's rows
    int *keep, *n_keep; short *kept_xy; // surviving point indices in visiting order, their rounded centres
    unsigned char *mask;                // [h][w]
    unsigned int *maxbits;              // masked maximum of eig (ordered-integer encoding)
    unsigned long long *cand; int cand_cap; int *n_cand;
    unsigned int *grid; int gw, gh, cell;
    float *corners; int corner_cap; int *n_corners;
};
struct DetRun {                         // per-launch parameters, by value
    const DetArgs *tasks;
    int img_sel, use_mask, use_base;        // use_base: setMask starts from DetArgs::base_mask (FISHEYE) instead of a white image
    int tracker_mode;                       // corners wanted = max_cnt - n_keep (feature_tracker.cpp:182,192); else max_corners (<= 0: no limit)
    int max_cnt, max_corners, corner_cap;   // corner_cap: capacity of the caller's output (<= DetArgs::corner_cap)
    double quality, min_dist;
};

#ifdef VIWB_HOST_EMU
static inline float det_mul(float a, float b) { return a * b; }
static inline float det_add(float a, float b) { return a + b; }
static inline float det_sub(float a, float b) { return a - b; }
static inline float det_sqrt(float a) { return sqrtf(a); }
static inline unsigned det_f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float det_u2f(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
#else
__device__ __forceinline__ float det_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float det_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float det_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float det_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ unsigned det_f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float det_u2f(unsigned u) { return __uint_as_float(u); }
#endif
// monotone float -> unsigned map (works for negative values too), so that maxima and sort keys are integer comparisons
VIWB_D unsigned det_encode(float f) { const unsigned u = det_f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
VIWB_D float det_decode(unsigned e) { return det_u2f((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e); }
VIWB_D int det_reflect(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// ------------------------------------------------------------------------------------------------------------------ setMask
// One warp per stream: rank the tracked points by track count (descending; equal counts keep their order -- std::sort leaves that
// unspecified), then walk them in that order: a point survives iff the mask is still 255 under its rounded position, i.e. it lies in
// no circle painted for an earlier survivor (feature_tracker.cpp:78-88).  Thread 0 also resets the per-tick counters of the stream.
VIWB_D void det_order_warp(const DetArgs &a, const DetRun &run, int tid, int nt, int *order) {
    int n = a.n_dev ? *a.n_dev : 0;
    if (n > DET_MAXPTS) n = DET_MAXPTS;
    if (tid == 0) { *a.maxbits = 0u; *a.n_cand = 0; *a.n_corners = 0; }
    for (int i = tid; i < n; i += nt) {
        const int ci = a.track_cnt[i];
        int rank = 0;
        for (int j = 0; j < n; j++) { const int cj = a.track_cnt[j]; rank += (cj > ci || (cj == ci && j < i)) ? 1 : 0; }
        order[rank] = i;
    }
    VIWB_SYNCWARP();
    int nk = 0;
    const int r = a.radius;
    for (int k = 0; k < n; k++) {
        const int i = order[k];
        const int x = (int)lrintf(a.pts[2 * i]), y = (int)lrintf(a.pts[2 * i + 1]);        // Point2f -> Point: cvRound
        bool hit = x < 0 || y < 0 || x >= a.w || y >= a.h;                                  // (inBorder() already guarantees this upstream)
        if (!hit && run.use_base && a.base_mask[(size_t)y * a.w + x] != 255) hit = true;
        for (int j = tid; j < nk && !hit; j += nt) {
            const int dy = abs(y - (int)a.kept_xy[2 * j + 1]), dx = abs(x - (int)a.kept_xy[2 * j]);
            if (dy <= r && dx <= (int)a.hw[dy]) hit = true;
        }
#ifndef VIWB_HOST_EMU
        hit = __any_sync(0xffffffffu, hit);
#endif
        if (!hit) {
            if (tid == 0) { a.kept_xy[2 * nk] = (short)x; a.kept_xy[2 * nk + 1] = (short)y; a.keep[nk] = i; }
            nk++;
        }
        VIWB_SYNCWARP();
    }
    if (tid == 0) *a.n_keep = nk;
}

// The mask image: a band of DET_BAND rows per block, white (or the fisheye mask) minus the filled circles of the survivors
// (cv::circle(mask, pt, MIN_DIST, 0, -1): midpoint circle, one run of 2*hw[|dy|]+1 pixels per row).
VIWB_D void det_mask_band(const DetArgs &a, const DetRun &run, int band, int tid, int nt, short *near_xy, int *n_near) {
    const int y0 = band * DET_BAND, y1 = (y0 + DET_BAND < a.h) ? y0 + DET_BAND : a.h, r = a.radius, nk = *a.n_keep;
    if (tid == 0) *n_near = 0;
    VIWB_SYNC();
    for (int j = tid; j < nk; j += nt) {                          // the circles that reach into this band (their order does not matter)
        const int cy = a.kept_xy[2 * j + 1];
        if (cy + r >= y0 && cy - r < y1) {
#ifdef VIWB_HOST_EMU
            const int m = (*n_near)++;
#else
            const int m = atomicAdd(n_near, 1);
#endif
            near_xy[2 * m] = a.kept_xy[2 * j]; near_xy[2 * m + 1] = (short)cy;
        }
    }
    VIWB_SYNC();
    // 16 pixels per thread and step: a circle is rejected for the whole group with two comparisons (a 61-pixel run meets few of the
    // 47 groups of a 752-pixel row), the survivors clear their bytes, one 16-byte store when the row start allows it
    const int m = *n_near, groups = (a.w + 15) / 16;
    for (int i = tid; i < (y1 - y0) * groups; i += nt) {
        const int y = y0 + i / groups, xb = (i % groups) * 16;
        const int nvalid = a.w - xb < 16 ? a.w - xb : 16;
        unsigned int px[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (run.use_base) {
#pragma unroll
            for (int k = 0; k < 16; k++) { const unsigned v = k < nvalid ? a.base_mask[(size_t)y * a.w + xb + k] : 255u; px[k >> 2] = (px[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (v << (8 * (k & 3))); }
        }
        unsigned clear = 0u;                                           // bit k: pixel xb + k lies in some circle
        for (int j = 0; j < m; j++) {
            const int dy = abs(y - (int)near_xy[2 * j + 1]);
            if (dy > r) continue;
            const int cx = near_xy[2 * j], half = a.hw[dy];
            int lo = cx - half - xb, hi = cx + half - xb;              // the run, in group coordinates
            if (hi < 0 || lo > 15) continue;
            lo = lo < 0 ? 0 : lo; hi = hi > 15 ? 15 : hi;
            clear |= ((2u << hi) - 1u) & ~((1u << lo) - 1u);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {                                  // four mask bits -> four mask bytes
            unsigned x = (clear >> (4 * q)) & 15u;
            x = (x | (x << 7) | (x << 14) | (x << 21)) & 0x01010101u;
            px[q] &= ~(x * 0xffu);
        }
        unsigned char *dst = a.mask + (size_t)y * a.w + xb;
#ifndef VIWB_HOST_EMU
        if (nvalid == 16 && ((size_t)dst & 15) == 0) { *reinterpret_cast<uint4 *>(dst) = make_uint4(px[0], px[1], px[2], px[3]); continue; }
#endif
        for (int k = 0; k < nvalid; k++) dst[k] = (unsigned char)(px[k >> 2] >> (8 * (k & 3)));
    }
}

// ------------------------------------------------------------------ cornerMinEigenVal + local maxima, one pass over the image
// A block owns a DET_OW x DET_OH patch of pixels.  It stages the source patch (+3 px halo, reflected), forms the derivative
// products over the patch +2 px in FP64, box-sums them with a sliding window (each thread walks 4 rows of one column: row sums
// (c0+c1)+c2, then (r0+r1)+r2, the order of a separable box filter), evaluates the eigenvalue over the patch +1 px, and then
//   * folds the masked maximum of its own pixels into DetArgs::maxbits,
//   * appends every 3x3 local maximum (non-zero, mask != 0, rows / columns 1..n-2) to the candidate list.
// The eigenvalue image never goes to HBM.  Whether a pixel is a 3x3 maximum of the thresholded image does not depend on the
// threshold once the pixel itself is above it, so the threshold (max * quality, known only after this pass) is applied to the
// candidate list by det_select_block.
enum { DET_OW = 62, DET_OH = 14, DET_EW = DET_OW + 2, DET_EH = DET_OH + 2, DET_CW = DET_OW + 4, DET_CH = DET_OH + 4, DET_SW = DET_OW + 6, DET_SH = DET_OH + 6,
       DET_NT = DET_EW * (DET_EH / 4) };                              // 64 columns x 4 row groups = 256 threads
VIWB_HD size_t det_tile_smem_bytes() { return (size_t)3 * DET_CH * DET_CW * 8 + (size_t)DET_SH * DET_SW * 4 + (size_t)DET_EH * DET_EW * 4; }
VIWB_D void det_commit_max(unsigned *dst, unsigned enc) {
#ifdef VIWB_HOST_EMU
    if (enc > *dst) *dst = enc;
#else
    enc = __reduce_max_sync(0xffffffffu, enc);
    if ((threadIdx.x & 31) == 0 && enc) atomicMax(dst, enc);
#endif
}
VIWB_D void det_append(const DetArgs &a, bool ok, unsigned long long key) {
#ifdef VIWB_HOST_EMU
    if (ok) { const int pos = (*a.n_cand)++; if (pos < a.cand_cap) a.cand[pos] = key; }
#else
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(a.n_cand, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (ok && pos < a.cand_cap) a.cand[pos] = key;
    }
#endif
}
// stages 1 and 2 of a tile; INTERIOR = the patch and its 3 px halo lie inside the image, so no coordinate is reflected
template <bool INTERIOR>
VIWB_D void det_tile_products(const DetArgs &a, const unsigned char *img, int x0, int y0, int tid, int nt, float *S, double *C0, double *C1, double *C2) {
    const int w = a.w, h = a.h;
    const double scale = 1.0 / (4 * 3) / 255.0;                   // 1 / (2^(ksize-1) * blockSize * 255)
    const float k1 = (float)scale, k2 = (float)(2.0 * scale);
    for (int i = tid; i < DET_SH * DET_SW; i += nt) {             // source patch, origin (x0-3, y0-3)
        const int rr = i / DET_SW, cc = i - rr * DET_SW;
        const int y = INTERIOR ? y0 - 3 + rr : det_reflect(y0 - 3 + rr, h), x = INTERIOR ? x0 - 3 + cc : det_reflect(x0 - 3 + cc, w);
        S[i] = (float)img[y * a.stride + x];
    }
    VIWB_SYNC();
    for (int i = tid; i < DET_CH * DET_CW; i += nt) {             // derivative products, origin (x0-2, y0-2)
        const int rr = i / DET_CW, cc = i - rr * DET_CW;
        // the box filter sees the product image reflected at the border: evaluate the derivatives at the reflected coordinate
        const int sr = INTERIOR ? rr + 1 : det_reflect(y0 - 2 + rr, h) - (y0 - 3), sc = INTERIOR ? cc + 1 : det_reflect(x0 - 2 + cc, w) - (x0 - 3);
        float xx = 0.f, xy = 0.f, yy = 0.f;
        if (INTERIOR || (sr >= 1 && sr <= DET_SH - 2 && sc >= 1 && sc <= DET_SW - 2)) {
            const float *p0 = S + (sr - 1) * DET_SW + sc, *p1 = p0 + DET_SW, *p2 = p1 + DET_SW;
            const float r0 = det_sub(p0[1], p0[-1]), r1 = det_sub(p1[1], p1[-1]), r2 = det_sub(p2[1], p2[-1]);
            const float dx = det_add(det_mul(k2, r1), det_mul(k1, det_add(r0, r2)));
            const float s0 = det_add(det_add(det_mul(k1, p0[-1]), det_mul(k2, p0[0])), det_mul(k1, p0[1]));
            const float s2 = det_add(det_add(det_mul(k1, p2[-1]), det_mul(k2, p2[0])), det_mul(k1, p2[1]));
            const float dy = det_sub(s2, s0);
            xx = det_mul(dx, dx); xy = det_mul(dx, dy); yy = det_mul(dy, dy);
        }
        C0[i] = (double)xx; C1[i] = (double)xy; C2[i] = (double)yy;
    }
    VIWB_SYNC();
}
VIWB_D void det_corner_tile(const DetArgs &a, const DetRun &run, int bx, int by, int tid, int nt, unsigned char *smem) {
    const int x0 = bx * DET_OW, y0 = by * DET_OH, w = a.w, h = a.h;
    const unsigned char *img = a.img[run.img_sel];
    double *C0 = (double *)smem, *C1 = C0 + DET_CH * DET_CW, *C2 = C1 + DET_CH * DET_CW;
    float *S = (float *)(C2 + DET_CH * DET_CW), *E = S + DET_SH * DET_SW;
    // the mask bytes of this thread's pixels are fetched first: their latency hides behind the whole tile computation
    const int rounds = (DET_OW * DET_OH + nt - 1) / nt;
    unsigned free_bits = 0u;
    for (int q = 0; q < rounds && q < 32; q++) {
        const int i = q * nt + tid, rr = i / DET_OW, cc = i - rr * DET_OW, y = y0 + rr, x = x0 + cc;
        if (i < DET_OW * DET_OH && y < h && x < w && (!run.use_mask || a.mask[(size_t)y * w + x])) free_bits |= 1u << q;
    }
    if (x0 >= 3 && y0 >= 3 && x0 + DET_OW + 3 <= w && y0 + DET_OH + 3 <= h) det_tile_products<true>(a, img, x0, y0, tid, nt, S, C0, C1, C2);
    else det_tile_products<false>(a, img, x0, y0, tid, nt, S, C0, C1, C2);
    for (int t = tid; t < DET_NT; t += nt) {                      // eigenvalues, origin (x0-1, y0-1): column c, rows 4g..4g+3
        const int c = t % DET_EW, g = t / DET_EW;
        double ra[3], rb[3], rc[3];                               // row sums of the three most recent product rows
        for (int k = 0; k < 6; k++) {
            const int j = (4 * g + k) * DET_CW + c;
            const double na = (C0[j] + C0[j + 1]) + C0[j + 2], nb = (C1[j] + C1[j + 1]) + C1[j + 2], nc = (C2[j] + C2[j + 1]) + C2[j + 2];
            if (k >= 2) {
                const float fa = det_mul((float)((ra[0] + ra[1]) + na), 0.5f), fb = (float)((rb[0] + rb[1]) + nb), fc = det_mul((float)((rc[0] + rc[1]) + nc), 0.5f);
                const float d = det_sub(fa, fc);
                E[(4 * g + k - 2) * DET_EW + c] = det_sub(det_add(fa, fc), det_sqrt(det_add(det_mul(d, d), det_mul(fb, fb))));
                ra[0] = ra[1]; rb[0] = rb[1]; rc[0] = rc[1];
                ra[1] = na; rb[1] = nb; rc[1] = nc;
            } else { ra[k] = na; rb[k] = nb; rc[k] = nc; }
        }
    }
    VIWB_SYNC();
    unsigned best = 0u;
    for (int q = 0; q < rounds; q++) {
        const int i = q * nt + tid;
        bool ok = false;
        unsigned long long key = 0ull;
        if (i < DET_OW * DET_OH) {
            const int rr = i / DET_OW, cc = i - rr * DET_OW, y = y0 + rr, x = x0 + cc;
            if (y < h && x < w) {
                const float *e = E + (rr + 1) * DET_EW + cc + 1;
                const float v = e[0];
                const bool free_px = q < 32 ? ((free_bits >> q) & 1u) != 0u : (!run.use_mask || a.mask[(size_t)y * w + x]);
                if (free_px) { const unsigned enc = det_encode(v); if (enc > best) best = enc; }
                if (free_px && v != 0.f && x >= 1 && y >= 1 && x <= w - 2 && y <= h - 2) {
                    ok = e[-DET_EW - 1] <= v && e[-DET_EW] <= v && e[-DET_EW + 1] <= v && e[-1] <= v && e[1] <= v && e[DET_EW - 1] <= v && e[DET_EW] <= v && e[DET_EW + 1] <= v;
                    key = ((unsigned long long)det_encode(v) << 32) | (unsigned)(y * w + x);
                }
            }
        }
        det_append(a, ok, key);
    }
    det_commit_max(a.maxbits, best);
}

VIWB_D float det_threshold(const DetArgs &a, const DetRun &run) {
    const unsigned mb = *a.maxbits;
    const float mx = mb ? det_decode(mb) : 0.f;                    // minMaxLoc over an empty mask leaves 0
    return (float)((double)mx * run.quality);
}

// ------------------------------------------------------------------------------------------------- sort + greedy selection
VIWB_D void det_bitonic_desc(unsigned long long *k, int P, int tid, int nt) {
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < P / 2; i += nt) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const unsigned long long x = k[lo], y = k[hi];
                if ((x < y) == ((lo & size) == 0)) { k[lo] = y; k[hi] = x; }
            }
            VIWB_SYNC();
        }
}
VIWB_D bool det_grid_hit(const unsigned *grid, int gw, int gh, int cell, int x, int y, double md2) {
    const int xc = x / cell, yc = y / cell;
    const int x1 = xc > 0 ? xc - 1 : 0, x2 = xc + 1 < gw ? xc + 1 : gw - 1, y1 = yc > 0 ? yc - 1 : 0, y2 = yc + 1 < gh ? yc + 1 : gh - 1;
    for (int yy = y1; yy <= y2; yy++) for (int xx = x1; xx <= x2; xx++) {
        const unsigned *c = grid + (size_t)(yy * gw + xx) * DET_SLOTS;
        for (int s = 0; s < DET_SLOTS; s++) {
            const unsigned p = c[s];
            if (p == 0xffffffffu) break;
            const int dx = x - (int)(p & 0xffffu), dy = y - (int)(p >> 16);
            if ((double)(dx * dx + dy * dy) < md2) return true;
        }
    }
    return false;
}
VIWB_D void det_grid_insert(unsigned *grid, int gw, int cell, int x, int y) {
    unsigned *c = grid + (size_t)((y / cell) * gw + x / cell) * DET_SLOTS;
    for (int s = 0; s < DET_SLOTS; s++) if (c[s] == 0xffffffffu) { c[s] = (unsigned)x | ((unsigned)y << 16); return; }
}
VIWB_HD int det_pow2_at_least(int n) { int p = 2; while (p < n) p <<= 1; return p; }
VIWB_HD size_t det_select_smem_bytes() { return (size_t)DET_SORT_SMEM * 8 + (size_t)DET_GRID_SMEM * DET_SLOTS * 4; }

VIWB_D void det_select_block(const DetArgs &a, const DetRun &run, int tid, int nt, unsigned char *smem, int *counter) {
    const int cap = (run.corner_cap > 0 && run.corner_cap < a.corner_cap) ? run.corner_cap : a.corner_cap;
    int want = run.tracker_mode ? run.max_cnt - *a.n_keep : (run.max_corners > 0 ? run.max_corners : cap);
    if (want > cap) want = cap;
    const int Mall = *a.n_cand;
    if (Mall > a.cand_cap) { if (tid == 0) *a.n_corners = -1; return; }      // candidate list overflow: reported, never truncated silently
    if (want <= 0 || Mall <= 0) { if (tid == 0) *a.n_corners = 0; return; }
    // threshold(eig, max * quality, THRESH_TOZERO): only candidates above it take part
    const float thr = det_threshold(a, run);
    if (tid == 0) { counter[0] = 0; counter[1] = 0; }
    VIWB_SYNC();
    int mine = 0;
    for (int i = tid; i < Mall; i += nt) mine += det_decode((unsigned)(a.cand[i] >> 32)) > thr ? 1 : 0;
#ifdef VIWB_HOST_EMU
    counter[0] += mine;
#else
    if (mine) atomicAdd(&counter[0], mine);
#endif
    VIWB_SYNC();
    const int M = counter[0];
    if (M <= 0) { if (tid == 0) *a.n_corners = 0; return; }
    unsigned long long *keys;
    int P;
    if (M <= DET_SORT_SMEM) {                                    // the usual case: the survivors are sorted in shared memory
        keys = (unsigned long long *)smem; P = det_pow2_at_least(M);
        for (int i = tid; i < Mall; i += nt) {
            const unsigned long long k = a.cand[i];
            if (det_decode((unsigned)(k >> 32)) > thr) {
#ifdef VIWB_HOST_EMU
                const int pos = counter[1]++;
#else
                const int pos = atomicAdd(&counter[1], 1);
#endif
                keys[pos] = k;
            }
        }
        for (int i = M + tid; i < P; i += nt) keys[i] = 0ull;
    } else {                                                     // sort the whole list in place; its first M entries are the survivors
        keys = a.cand; P = det_pow2_at_least(Mall);
        for (int i = Mall + tid; i < P; i += nt) keys[i] = 0ull;
    }
    const int cells = a.gw * a.gh;
    unsigned *grid = (cells <= DET_GRID_SMEM) ? (unsigned *)(smem + (size_t)DET_SORT_SMEM * 8) : a.grid;
    for (int i = tid; i < cells * DET_SLOTS; i += nt) grid[i] = 0xffffffffu;
    VIWB_SYNC();
    det_bitonic_desc(keys, P, tid, nt);
    const bool use_grid = run.min_dist >= 1.0;
    const double md2 = run.min_dist * run.min_dist;
    const int w = a.w;
    int count = 0;
#ifdef VIWB_HOST_EMU
    for (int i = 0; i < M && count < want; i++) {
        const unsigned ofs = (unsigned)keys[i];
        const int y = ofs / w, x = ofs - y * w;
        if (use_grid) { if (det_grid_hit(grid, a.gw, a.gh, a.cell, x, y, md2)) continue; det_grid_insert(grid, a.gw, a.cell, x, y); }
        a.corners[2 * count] = (float)x; a.corners[2 * count + 1] = (float)y; count++;
    }
    *a.n_corners = count;
#else
    if (tid >= 32) return;
    const int lane = tid;
    // 32 candidates at a time: each lane screens its candidate against the corners accepted so far; the survivors are then resolved
    // in rank order inside the warp (the best one is accepted, the others re-test against it) -- the same result as the serial loop
    for (int base = 0; base < M && count < want; base += 32) {
        const int idx = base + lane;
        const unsigned ofs = idx < M ? (unsigned)keys[idx] : 0u;
        const int y = ofs / w, x = ofs - y * w;
        bool alive = idx < M && !(use_grid && det_grid_hit(grid, a.gw, a.gh, a.cell, x, y, md2));
        unsigned m = __ballot_sync(0xffffffffu, alive);
        while (m && count < want) {
            const int l = __ffs(m) - 1;
            const int sx = __shfl_sync(0xffffffffu, x, l), sy = __shfl_sync(0xffffffffu, y, l);
            if (lane == l) {
                if (use_grid) det_grid_insert(grid, a.gw, a.cell, x, y);
                a.corners[2 * count] = (float)x; a.corners[2 * count + 1] = (float)y;
                alive = false;
            } else if (alive && use_grid) {
                const int dx = x - sx, dy = y - sy;
                if ((double)(dx * dx + dy * dy) < md2) alive = false;
            }
            count++;
            m = __ballot_sync(0xffffffffu, alive);
        }
        __syncwarp();
    }
    if (lane == 0) *a.n_corners = count;
#endif
}

#ifndef VIWB_HOST_EMU
__global__ void __launch_bounds__(32) det_order_kernel(DetRun run) {
    __shared__ int order[DET_MAXPTS];
    det_order_warp(run.tasks[blockIdx.x], run, threadIdx.x, 32, order);
}
__global__ void __launch_bounds__(256) det_mask_kernel(DetRun run) {
    __shared__ short near_xy[2 * DET_MAXPTS];
    __shared__ int n_near;
    det_mask_band(run.tasks[blockIdx.y], run, blockIdx.x, threadIdx.x, blockDim.x, near_xy, &n_near);
}
__global__ void __launch_bounds__(DET_NT) det_corners_kernel(DetRun run, int tiles_x) {
    extern __shared__ __align__(16) unsigned char det_sm[];
    det_corner_tile(run.tasks[blockIdx.y], run, blockIdx.x % tiles_x, blockIdx.x / tiles_x, threadIdx.x, blockDim.x, det_sm);
}
__global__ void __launch_bounds__(1024) det_select_kernel(DetRun run) {
    extern __shared__ __align__(16) unsigned char det_sel_sm[];
    __shared__ int counter[2];
    det_select_block(run.tasks[blockIdx.x], run, threadIdx.x, blockDim.x, det_sel_sm, counter);
}
#endif

}  // namespace viwb

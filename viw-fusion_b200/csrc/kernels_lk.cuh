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
    int trowI[LK_MAXLVL], trowJ[LK_MAXLVL];      // first row of the template / search image inside the level's stacked tensor (TMA coordinates)
    int tma;                  // the images live in an LK batch's level stacks: interior windows are staged by TMA tile loads
};

VIWB_HD int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
VIWB_HD int lk_pix(const LkImage &im, int l, int x, int y) { return im.img[l][(size_t)reflect101(y, im.h[l]) * im.stride[l] + reflect101(x, im.w[l])]; }
VIWB_HD int cv_round_f(float v) { return (int)lrintf(v); }                 // cvRound: round half to even
VIWB_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }   // CV_DESCALE

// ---- pyrDown: dst (dw x dh) from src (sw x sh).  One work item = a 4 x 4 block of outputs: 11 source rows x 4 aligned 32-bit loads,
// the horizontal [1 4 6 4 1] sums of a row (two funnel shifts + four byte dot products) computed once and added into the (up to three)
// output rows that use it -- the same integers as the 2-D stencil at 12 instead of 32 instructions per output.  Blocks touching the
// left / right border, and the tails of images whose size is not a multiple of 4, gather their bytes through reflected indices
// (pyr_down_block_edge) and are numbered AFTER all interior blocks, so they share no warp with them.
#ifndef PYR_BATCH
#define PYR_BATCH 6
#endif
struct PyrArgs { const uint8_t *src; uint8_t *dst; int sw, sh, sstride, dw, dh, dstride; };
VIWB_HD int pyr_items_wh(int dw, int dh) { return ((dw + 3) / 4) * ((dh + 3) / 4); }
VIWB_HD int pyr_items(const PyrArgs &a) { return pyr_items_wh(a.dw, a.dh); }
// one interior 4 x 4 block; INNER = none of the 11 source rows needs reflecting (straight-line code, no per-row branches)
template <bool INNER>
VIWB_D void pyr_down_block(const PyrArgs &a, int x, int y) {
    int acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[j][k] = 0;
    // the 11 rows are fetched in batches of PYR_BATCH rows: all loads of a batch are in flight before the first one is consumed
    // (a thread that consumes row by row pays the DRAM latency eleven times in a row and the kernel runs latency-bound)
#pragma unroll
    for (int r0 = 0; r0 < 11; r0 += PYR_BATCH) {
        uint32_t wd[PYR_BATCH][4];
#pragma unroll
        for (int q = 0; q < PYR_BATCH; q++) {
            const int r = r0 + q;
            if (r < 11) {
                const int ry = INNER ? 2 * y - 2 + r : reflect101(2 * y - 2 + r, a.sh);
                const uint32_t *row = reinterpret_cast<const uint32_t *>(a.src + (size_t)ry * a.sstride + 2 * x - 4);
#if defined(PYR_ASM_LOADS) && !defined(VIWB_HOST_EMU)
                // volatile: the loads of a batch keep their program order and are not sunk next to their uses
                asm volatile("ld.global.nc.u32 %0, [%4];\n\tld.global.nc.u32 %1, [%4+4];\n\tld.global.nc.u32 %2, [%4+8];\n\tld.global.nc.u32 %3, [%4+12];"
                             : "=r"(wd[q][0]), "=r"(wd[q][1]), "=r"(wd[q][2]), "=r"(wd[q][3]) : "l"(row));
#else
                wd[q][0] = row[0]; wd[q][1] = row[1]; wd[q][2] = row[2]; wd[q][3] = row[3];
#endif
            }
        }
#pragma unroll
        for (int q = 0; q < PYR_BATCH; q++) {
            const int r = r0 + q;                                      // source row 2y - 2 + r feeds output row j with weight wgt[r - 2j]
            if (r < 11) {
                const uint32_t w0 = wd[q][0], w1 = wd[q][1], w2 = wd[q][2], w3 = wd[q][3];
                // bytes b[0..15] = source columns 2x-4 .. 2x+11 (b[0] = low byte of w0); output k (0..3) is centred on b[4 + 2k]:
                // taps b[2+2k .. 5+2k] are one (funnel-shifted) 32-bit group for a 4-way byte dot product with (1, 4, 6, 4), plus b[6+2k]
                int t[4];
#ifdef VIWB_HOST_EMU
                int b[16];
                for (int k = 0; k < 4; k++) { b[k] = (w0 >> (8 * k)) & 0xff; b[4 + k] = (w1 >> (8 * k)) & 0xff; b[8 + k] = (w2 >> (8 * k)) & 0xff; b[12 + k] = (w3 >> (8 * k)) & 0xff; }
                for (int k = 0; k < 4; k++) { const int c = 4 + 2 * k; t[k] = b[c - 2] + b[c + 2] + 4 * (b[c - 1] + b[c + 1]) + 6 * b[c]; }
#else
                const uint32_t g0 = __funnelshift_r(w0, w1, 16), g2 = __funnelshift_r(w1, w2, 16);
                t[0] = (int)__dp4a(g0, 0x04060401u, (w1 >> 16) & 0xffu); t[1] = (int)__dp4a(w1, 0x04060401u, w2 & 0xffu);
                t[2] = (int)__dp4a(g2, 0x04060401u, (w2 >> 16) & 0xffu); t[3] = (int)__dp4a(w2, 0x04060401u, w3 & 0xffu);
#endif
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int tap = r - 2 * j;                         // compile-time after unrolling
                    if (tap >= 0 && tap <= 4) {
                        const int wv = tap == 0 || tap == 4 ? 1 : (tap == 2 ? 6 : 4);
#pragma unroll
                        for (int k = 0; k < 4; k++) acc[j][k] += wv * t[k];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) o |= (uint32_t)((acc[j][k] + 128) >> 8) << (8 * k);
        *reinterpret_cast<uint32_t *>(a.dst + (size_t)(y + j) * a.dstride + x) = o;
    }
}
// a 4 x 4 block at the image border (left / right columns, tails of sizes that are not multiples of 4): the 16 source columns are
// reflected once per block and the rows once per row, the bytes are gathered into the same four words, and the arithmetic is the
// interior one -- about 4x an interior block instead of 16 independent 25-tap pixels
VIWB_D void pyr_down_block_edge(const PyrArgs &a, int x, int y) {
    int cxi[16];
#pragma unroll
    for (int i = 0; i < 16; i++) cxi[i] = reflect101(2 * x - 4 + i, a.sw);
    int acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[j][k] = 0;
#pragma unroll 1
    for (int r = 0; r < 11; r++) {
        const uint8_t *row = a.src + (size_t)reflect101(2 * y - 2 + r, a.sh) * a.sstride;
        int b[16];
#pragma unroll
        for (int i = 2; i < 13; i++) b[i] = row[cxi[i]];            // outputs k = 0..3 use bytes 2 .. 12
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = 4 + 2 * k, t = b[c - 2] + b[c + 2] + 4 * (b[c - 1] + b[c + 1]) + 6 * b[c];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int tap = r - 2 * j;
                const int wv = (tap == 0 || tap == 4) ? 1 : (tap == 2 ? 6 : ((tap == 1 || tap == 3) ? 4 : 0));
                acc[j][k] += wv * t;
            }
        }
    }
    for (int j = 0; j < 4 && y + j < a.dh; j++)
        for (int k = 0; k < 4 && x + k < a.dw; k++) a.dst[(size_t)(y + j) * a.dstride + x + k] = (uint8_t)((acc[j][k] + 128) >> 8);
}
// Item order: first the blocks whose columns are interior (block columns 1 .. ci), row by row, then the border columns -- so that the
// few border blocks fill warps of their own instead of stalling one lane in most warps (with 94 blocks per row and one slow lane per
// 47, two warps in three used to wait for a border block).
VIWB_D void pyr_down_item(const PyrArgs &a, int idx) {
    const int sx = (a.dw + 3) / 4, sy = (a.dh + 3) / 4;
    if (idx >= sx * sy) return;
    int xmax = (a.sw - 12) / 2; if (a.dw - 4 < xmax) xmax = a.dw - 4;     // interior block: x >= 2, 2x + 12 <= sw, x + 4 <= dw
    int ci = xmax >= 4 ? xmax / 4 : 0; if (ci > sx - 1) ci = sx - 1;
    const int ce = sx - ci;
    int bx, by;
    if (idx < ci * sy) { by = idx / ci; bx = 1 + idx - by * ci; }
    else { const int e = idx - ci * sy; by = e / ce; const int k = e - by * ce; bx = k == 0 ? 0 : ci + k; }
    const int x = 4 * bx, y = 4 * by;
    if (bx >= 1 && bx <= ci && y + 4 <= a.dh) {
        if (2 * y - 2 >= 0 && 2 * y + 8 < a.sh) pyr_down_block<true>(a, x, y);
        else pyr_down_block<false>(a, x, y);
    } else pyr_down_block_edge(a, x, y);
}

// ---- the tracker: one warp per point, no block-level barriers.
// Work items of a warp: the 21 window rows x 3 segments of 7 pixels = 63 items, two per lane, so that neighbouring samples
// share their loads and every shared-memory offset inside an item is a compile-time constant.  Per level the warp
//   1. stages the source patch: ONE TMA tile load (cp.async.bulk.tensor.2d, 32 x 32 bytes) when the 24 x 24 patch lies inside the image,
//      the reflect-101 byte / word path otherwise,
//   2. interpolates the 23 x 23 intensities once (two byte-pair dot products per sample, dp2a) and derives template + Scharr
//      gradients from them (zero-padded Scharr like OpenCV's constant-border derivative buffer on the border path),
//   3. keeps its 14 template samples in registers as c0 = 256 - 512 * I (so that one arithmetic shift finishes a residual) with the
//      gradients unpacked,
//   4. stages a 32 x 32 region of the search image around the estimate (TMA tile when inside the image) and samples it in the
//      Gauss-Newton iterations: per row three aligned words, two byte permutes to undo the misalignment, two byte-pair dot products
//      per sample; restaged only if the estimate leaves it,
//   5. reduces the integer products with the warp-wide integer adder (redux.sync): exact, no shuffle chain.
#ifdef VIWB_HOST_EMU
enum { LK_W = 1 };
#else
enum { LK_W = 32 };
#endif
#ifndef LK_MINB
#define LK_MINB 5
#endif
enum { LK_ITEMS = 63, LK_IPL = (LK_ITEMS + LK_W - 1) / LK_W, LK_PPB = 4,         // items, items per lane, points per block
       LK_PS = 48,                                                             // byte stride of the staged source patch (TMA box row: 48 bytes, its origin 16-byte aligned)
       LK_SLACK = 4, LK_JROWS = 32, LK_JS = 48,                                // staged search region: 32 rows x 48 bytes (TMA box)
       LK_BS = 23, LK_DBYTES = 2208,                                           // interpolated-intensity image 23 x 23 ints (shares the Scharr sample buffer)
       LK_OFF_J = 32 * LK_PS, LK_OFF_D = LK_OFF_J + LK_JROWS * LK_JS, LK_OFF_BAR = LK_OFF_D + LK_DBYTES,
       LK_WARP_SMEM = 5376 };                                                  // 1536 + 1536 + 2208 + 16 (mbarriers) rounded up to a multiple of 128
VIWB_HD size_t lk_smem_bytes(int warps) { return (size_t)warps * LK_WARP_SMEM; }

// exact sum over the warp of one int32 per lane (every lane gets it)
VIWB_D long long lk_warp_sum(int p) {
#ifdef VIWB_HOST_EMU
    return (long long)p;
#else
    const int hi = __reduce_add_sync(0xffffffffu, p >> 12), lo = __reduce_add_sync(0xffffffffu, p & 0xfff);
    return ((long long)hi << 12) + (long long)lo;
#endif
}
VIWB_HD bool lk_outside(int ix, int iy, int cols, int rows) { return ix < -LK_WIN || ix >= cols || iy < -LK_WIN || iy >= rows; }

// rows [y0, y0+nrows) x bytes [x0a, x0a + 4*nwords) of an image into a byte buffer with row stride 4*nwords.
// x0a is a multiple of 4; interior regions use aligned 32-bit loads, the rest resolves reflect-101 byte by byte.
VIWB_D void lk_stage(uint8_t *buf, const uint8_t *img, int stride, int cols, int rows, int x0a, int y0, int nrows, int nwords, int lane) {
    VIWB_SYNCWARP();
    if (x0a >= 0 && x0a + 4 * nwords <= cols) {          // columns inside the image: aligned words; only the row index may need reflecting
        const int total = nrows * nwords;
        const bool rows_in = y0 >= 0 && y0 + nrows <= rows;
        for (int e = lane; e < total; e += LK_W) {
            const int y = e / nwords, x = e - y * nwords;
            const int sy = rows_in ? y0 + y : reflect101(y0 + y, rows);
            reinterpret_cast<uint32_t *>(buf)[e] = *reinterpret_cast<const uint32_t *>(img + (size_t)sy * stride + x0a + 4 * x);
        }
    } else {        // left / right border: still one aligned word per item wherever the word lies inside the row; only the overhanging words gather
                    // their four bytes through reflected columns (a lane per column with byte loads row after row measured 17 % of lk_track: r02r capture)
        const int total = nrows * nwords;
        for (int e = lane; e < total; e += LK_W) {
            const int y = e / nwords, x = e - y * nwords, xc = x0a + 4 * x;
            const uint8_t *src = img + (size_t)reflect101(y0 + y, rows) * stride;
            uint32_t wv;
            if (xc >= 0 && xc + 4 <= cols) wv = *reinterpret_cast<const uint32_t *>(src + xc);
            else wv = (uint32_t)src[reflect101(xc, cols)] | ((uint32_t)src[reflect101(xc + 1, cols)] << 8) | ((uint32_t)src[reflect101(xc + 2, cols)] << 16) | ((uint32_t)src[reflect101(xc + 3, cols)] << 24);
            reinterpret_cast<uint32_t *>(buf)[e] = wv;
        }
    }
    VIWB_SYNCWARP();
}

// ---- TMA staging (sm_100a): per pyramid level ONE tensor map over the level's stacked images [slots * streams * rows][width] (u8, row pitch a
// multiple of 16 bytes), box 48 bytes x 32 rows.  The unit wants the tile's first byte 16-byte aligned (a tile at a byte column that is not a
// multiple of 16 raises "illegal instruction", profiles/r02i_tma_probe.txt: the probes at x = 8 bytes fault, those at x = 32 bytes load), so
// the tile starts at the aligned column at or below the window and is 48 bytes wide (23 needed + 15 of misalignment + slack).  A warp's
// lane 0 arms the warp's mbarrier with the tile's byte count and issues the copy, the warp waits on the barrier's phase.  Columns right of the
// image are zero-filled by the unit; only windows whose needed pixels are inside the image take this path, and samples are only taken from the
// part of a tile that lies inside the image (rows below it belong to the next image of the stack).
struct alignas(64) LkMaps { unsigned long long opaque[LK_MAXLVL][16]; };      // LK_MAXLVL x CUtensorMap (128 bytes each; the descriptor must sit 64-byte aligned, also in kernel parameter space), encoded by the host
#ifndef VIWB_HOST_EMU
VIWB_D unsigned lk_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
VIWB_D void lk_bar_init(unsigned long long *bar, int lane) {      // bar[0]: source-patch copies, bar[1]: search-region copies
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lk_smem_addr(bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lk_smem_addr(bar + 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the TMA unit (async proxy) must see the initialised barriers
    }
    __syncwarp();
}
// issue: tile (x, y) .. (x + 47, y + 31) of the level's stack (x a multiple of 16) into dst (128-byte aligned), completion on `bar`; returns at once
VIWB_D void lk_tma_issue(uint8_t *dst, const void *map, int x, int y, unsigned long long *bar, int lane) {
    __syncwarp();                                  // every lane is done reading the buffer that is about to be overwritten
    if (lane == 0) {
        const unsigned b = lk_smem_addr(bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(LK_JS * LK_JROWS) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(lk_smem_addr(dst)), "l"(map), "r"(x), "r"(y), "r"(b) : "memory");
    }
}
// wait: the whole warp spins on the barrier's phase; phase = this barrier's count of finished copies & 1
VIWB_D void lk_tma_wait(unsigned long long *bar, unsigned &phase) {
    const unsigned b = lk_smem_addr(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(phase) : "memory");
    }
    phase ^= 1u;
}
#endif

#ifndef VIWB_HOST_EMU
// c + (signed 16-bit halves of w) . (unsigned bytes 0,1 / 2,3 of px)
VIWB_D int lk_dp2a_lo(unsigned w, unsigned px, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c)); return d; }
VIWB_D int lk_dp2a_hi(unsigned w, unsigned px, int c) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c)); return d; }
#endif
// bilinear samples of 7 consecutive pixels from two staged rows (row stride LK_JS = LK_PS bytes) at byte offset xoff (any alignment):
// out[k] = c[k] + top[k] w00 + top[k+1] w01 + bot[k] w10 + bot[k+1] w11, wt = w00 | w01 << 16, wb = w10 | w11 << 16 as SIGNED 16-bit halves
// (|w| <= 2^14; w11 = 2^14 - w00 - w01 - w10 can come out as -1, lkpyramid.cpp computes it the same way)
VIWB_D void lk_sample7(const uint8_t *rowbase, int xoff, unsigned wt, unsigned wb, const int *c, int *out) {
#ifdef VIWB_HOST_EMU
    const uint8_t *q = rowbase + xoff;
    const int w00 = (int)(short)(wt & 0xffffu), w01 = (int)(short)(wt >> 16), w10 = (int)(short)(wb & 0xffffu), w11 = (int)(short)(wb >> 16);
    for (int k = 0; k < 7; k++) out[k] = c[k] + q[k] * w00 + q[k + 1] * w01 + q[k + LK_JS] * w10 + q[k + LK_JS + 1] * w11;
#else
    const int mis = xoff & 3;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(rowbase + (xoff - mis));
    const unsigned sel = 0x3210u + 0x1111u * (unsigned)mis;
    const uint32_t t0 = wp[0], t1 = wp[1], t2 = wp[2], b0 = wp[LK_JS / 4], b1 = wp[LK_JS / 4 + 1], b2 = wp[LK_JS / 4 + 2];
    const unsigned ta = __byte_perm(t0, t1, sel), tb = __byte_perm(t1, t2, sel);      // bytes 0..3, 4..7 of the top row
    const unsigned ba = __byte_perm(b0, b1, sel), bb = __byte_perm(b1, b2, sel);
    const unsigned ta1 = __byte_perm(ta, tb, 0x4321), tb1 = tb >> 8, ba1 = __byte_perm(ba, bb, 0x4321), bb1 = bb >> 8;      // bytes 1..4, 5..7
    out[0] = lk_dp2a_lo(wb, ba, lk_dp2a_lo(wt, ta, c[0]));
    out[1] = lk_dp2a_lo(wb, ba1, lk_dp2a_lo(wt, ta1, c[1]));
    out[2] = lk_dp2a_hi(wb, ba, lk_dp2a_hi(wt, ta, c[2]));
    out[3] = lk_dp2a_hi(wb, ba1, lk_dp2a_hi(wt, ta1, c[3]));
    out[4] = lk_dp2a_lo(wb, bb, lk_dp2a_lo(wt, tb, c[4]));
    out[5] = lk_dp2a_lo(wb, bb1, lk_dp2a_lo(wt, tb1, c[5]));
    out[6] = lk_dp2a_hi(wb, bb, lk_dp2a_hi(wt, tb, c[6]));
#endif
}

VIWB_D void lk_track_warp(const LkArgs &a, const LkMaps *maps, int pt, int lane, unsigned char *smem_raw) {
    uint8_t *pbuf = (uint8_t *)smem_raw;                      // 32 rows x 48 bytes (24 rows used)
    uint8_t *jbuf = smem_raw + LK_OFF_J;                      // 32 rows x 48 bytes
    short *dpatch = (short *)(smem_raw + LK_OFF_D);           // 22 x 22 x (dx, dy)
    int *Bimg = (int *)dpatch;                                // or: 23 x 23 interpolated intensities (interior patches)
    const int npts = a.n_dev ? *a.n_dev : a.n;
    if (pt >= npts) return;
#ifndef VIWB_HOST_EMU
    unsigned long long *bar = (unsigned long long *)(smem_raw + LK_OFF_BAR);
    unsigned phaseP = 0, phaseJ = 0;
    bool p_pending = false, j_pending = false;      // a source-patch / search-region copy is in flight
    const bool tma = maps != nullptr && a.tma != 0;
    if (tma) lk_bar_init(bar, lane);
#else
    (void)maps;
#endif
    const int max_level = a.max_level, max_iter = a.max_iter, flags = a.flags;
    const float eps2 = a.eps2, min_eig = a.min_eig;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float px0 = a.prev_pts[2 * pt], py0 = a.prev_pts[2 * pt + 1];
    bool status = true; float errv = 0.f;
    float npx = 0.f, npy = 0.f;     // nextPts[ptidx] (window centre coordinates)
    int Ic[LK_IPL][7], gx[LK_IPL][7], gy[LK_IPL][7];    // 256 - 512 * template intensity (5 fractional bits), template gradients of the lane's samples
    int irow[LK_IPL], icol[LK_IPL];
#pragma unroll
    for (int s = 0; s < LK_IPL; s++) { const int it = lane + LK_W * s; irow[s] = it / 3; icol[s] = 7 * (it - 3 * irow[s]); }
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
        // 1. source patch rows [ipy-1, ipy+22]; pixel (ipx-1+x) of a row sits at byte psx + x.  The search region of this level is requested
        //    first (its position is known), so that its latency hides behind the template construction; the source patch itself was
        //    requested while the previous level iterated (its position only depends on the tracked point).
        int psx;
        const int jc = a.J.w[level], jr = a.J.h[level], jstr = a.J.stride[level];
        const uint8_t *jimg = a.J.img[level];
        int jx0 = 0, jy0 = 0;          // origin of the staged search region
        int jx1 = 0, jy1 = 0;          // end of its usable part (a TMA tile may hang over the image's right / bottom edge; the reflect path fills its whole box)
        bool staged = false;
#ifndef VIWB_HOST_EMU
        // TMA tile that holds the 23 x 23 samples at (inx, iny), which must lie inside the image: origin a multiple of 16 bytes in x, 4 .. 19 left of inx
        auto tile_j = [&](int inx, int iny) {
            jx0 = (inx - 4) & ~15; if (jx0 < 0) jx0 = 0;
            jy0 = iny - 4; if (jy0 > jr - LK_JROWS) jy0 = jr - LK_JROWS; if (jy0 < 0) jy0 = 0;
            jx1 = jx0 + LK_JS < jc ? jx0 + LK_JS : jc; jy1 = jy0 + LK_JROWS < jr ? jy0 + LK_JROWS : jr;
            lk_tma_issue(jbuf, &maps->opaque[level][0], jx0, a.trowJ[level] + jy0, bar + 1, lane);
        };
        if (tma) {
            const int inx = (int)floorf(nx - (float)LK_HALF), iny = (int)floorf(ny - (float)LK_HALF);
            if (inx >= 0 && iny >= 0 && inx + 23 <= jc && iny + 23 <= jr) { tile_j(inx, iny); j_pending = true; staged = true; }
        }
        if (tma && ipx >= 1 && ipy >= 1 && ipx + 23 <= cols && ipy + 23 <= rows) {
            const int pxa = (ipx - 1) & ~15;
            if (!p_pending) lk_tma_issue(pbuf, &maps->opaque[level][0], pxa, a.trowI[level] + ipy - 1, bar, lane);
            lk_tma_wait(bar, phaseP); p_pending = false;
            psx = (ipx - 1) - pxa;
        } else
#endif
        {
            const int pxa = (ipx - 1) & ~3;
            psx = (ipx - 1) - pxa;
            lk_stage(pbuf, a.I.img[level], a.I.stride[level], cols, rows, pxa, ipy - 1, 24, LK_PS / 4, lane);
        }
        const float fa = ppx - ipx, fb = ppy - ipy;
        const int iw00 = cv_round_f((1.f - fa) * (1.f - fb) * (1 << 14)), iw01 = cv_round_f(fa * (1.f - fb) * (1 << 14));
        const int iw10 = cv_round_f((1.f - fa) * fb * (1 << 14)), iw11 = (1 << 14) - iw00 - iw01 - iw10;
        int sA11 = 0, sA12 = 0, sA22 = 0;              // per lane < 14 * 2^24
        if (ipx >= 0 && ipy >= 0 && ipx + 22 <= cols && ipy + 22 <= rows) {
            // 2a/3a. interior patch: the bilinear weights commute with the (integer, unrounded) Scharr stencil, so interpolate the
            //        intensities once, B(y,x) = sum_i w_i I(.), and take the stencil of B -- the same integers as interpolating the
            //        four Scharr samples, with a third of the work.  Items = 23 rows x 4 groups of 7 (the last group has 2 columns).
            {
                const unsigned wt = ((unsigned)iw00 & 0xffffu) | ((unsigned)iw01 << 16), wb = ((unsigned)iw10 & 0xffffu) | ((unsigned)iw11 << 16);
                const int zero7[7] = {0, 0, 0, 0, 0, 0, 0};
                for (int e = lane; e < LK_BS * 4; e += LK_W) {
                    const int y = e >> 2, g = e & 3;
                    int v[7];
                    lk_sample7(pbuf + y * LK_PS, psx + 7 * g, wt, wb, zero7, v);      // (bytes past column 23 belong to the staged row: read, not used)
                    const int n = g < 3 ? 7 : 2;
                    for (int k = 0; k < 7; k++) if (k < n) Bimg[y * LK_BS + 7 * g + k] = v[k];
                }
            }
            VIWB_SYNCWARP();
#pragma unroll
            for (int s = 0; s < LK_IPL; s++) {
                if (lane + LK_W * s < LK_ITEMS) {
                    const int *b0 = Bimg + irow[s] * LK_BS + icol[s], *b1 = b0 + LK_BS, *b2 = b1 + LK_BS;
#pragma unroll
                    for (int j = 0; j < 7; j++) {
                        const int ival = descale(b1[j + 1], 9);
                        const int ix = descale((b0[j + 2] + b2[j + 2]) * 3 + b1[j + 2] * 10 - ((b0[j] + b2[j]) * 3 + b1[j] * 10), 14);
                        const int iy = descale(((b2[j + 2] - b0[j + 2]) + (b2[j] - b0[j])) * 3 + (b2[j + 1] - b0[j + 1]) * 10, 14);
                        Ic[s][j] = 256 - (ival << 9); gx[s][j] = ix; gy[s][j] = iy;
                        sA11 += ix * ix; sA12 += ix * iy; sA22 += iy * iy;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 7; j++) { Ic[s][j] = 256; gx[s][j] = 0; gy[s][j] = 0; }
                }
            }
        } else {
            // 2b. Scharr samples at [ipy, ipy+21] x [ipx, ipx+21]; zero outside the image
            for (int e = lane; e < LK_DPATCH * LK_DPATCH; e += LK_W) {
                const int y = e / LK_DPATCH, x = e - y * LK_DPATCH, gxx = ipx + x, gyy = ipy + y;
                int ddx = 0, ddy = 0;
                if (gxx >= 0 && gyy >= 0 && gxx < cols && gyy < rows) {
                    const uint8_t *p0 = pbuf + y * LK_PS + psx + x, *p1 = p0 + LK_PS, *p2 = p1 + LK_PS;
                    ddx = (p0[2] + p2[2]) * 3 + p1[2] * 10 - ((p0[0] + p2[0]) * 3 + p1[0] * 10);
                    ddy = ((p2[2] - p0[2]) + (p2[0] - p0[0])) * 3 + (p2[1] - p0[1]) * 10;
                }
                dpatch[2 * e] = (short)ddx; dpatch[2 * e + 1] = (short)ddy;
            }
            VIWB_SYNCWARP();
            // 3b. template
#pragma unroll
            for (int s = 0; s < LK_IPL; s++) {
                if (lane + LK_W * s < LK_ITEMS) {
                    const uint8_t *p = pbuf + (irow[s] + 1) * LK_PS + psx + icol[s] + 1;
                    const short *d = dpatch + 2 * (irow[s] * LK_DPATCH + icol[s]);
#pragma unroll
                    for (int j = 0; j < 7; j++) {
                        const int ival = descale(p[j] * iw00 + p[j + 1] * iw01 + p[j + LK_PS] * iw10 + p[j + LK_PS + 1] * iw11, 9);
                        const int ix = descale(d[2 * j] * iw00 + d[2 * j + 2] * iw01 + d[2 * j + 2 * LK_DPATCH] * iw10 + d[2 * j + 2 * LK_DPATCH + 2] * iw11, 14);
                        const int iy = descale(d[2 * j + 1] * iw00 + d[2 * j + 3] * iw01 + d[2 * j + 2 * LK_DPATCH + 1] * iw10 + d[2 * j + 2 * LK_DPATCH + 3] * iw11, 14);
                        Ic[s][j] = 256 - (ival << 9); gx[s][j] = ix; gy[s][j] = iy;
                        sA11 += ix * ix; sA12 += ix * iy; sA22 += iy * iy;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 7; j++) { Ic[s][j] = 256; gx[s][j] = 0; gy[s][j] = 0; }
                }
            }
        }
#ifndef VIWB_HOST_EMU
        if (tma && level > 0) {      // the template is in registers: the source-patch buffer is free for the next level's patch
            const float sc2 = (float)(1. / (1 << (level - 1)));
            const int qx = (int)floorf(px0 * sc2 - (float)LK_HALF), qy = (int)floorf(py0 * sc2 - (float)LK_HALF);
            if (qx >= 1 && qy >= 1 && qx + 23 <= a.I.w[level - 1] && qy + 23 <= a.I.h[level - 1]) {
                lk_tma_issue(pbuf, &maps->opaque[level - 1][0], (qx - 1) & ~15, a.trowI[level - 1] + qy - 1, bar, lane);
                p_pending = true;
            }
        }
#endif
        const long long A11 = lk_warp_sum(sA11), A12 = lk_warp_sum(sA12), A22 = lk_warp_sum(sA22);
        const float fA11 = (float)A11 * FLT_SCALE, fA12 = (float)A12 * FLT_SCALE, fA22 = (float)A22 * FLT_SCALE;
        float D = fA11 * fA22 - fA12 * fA12;
        const float minEig = (fA22 + fA11 - sqrtf((fA11 - fA22) * (fA11 - fA22) + 4.f * fA12 * fA12)) / (2 * LK_WIN * LK_WIN);
        if (minEig < min_eig || D < 1.1920929e-07f) {
#ifndef VIWB_HOST_EMU
            if (j_pending) { lk_tma_wait(bar + 1, phaseJ); j_pending = false; }
#endif
            if (level == 0) status = false;
            continue;
        }
        D = 1.f / D;
        nx -= (float)LK_HALF; ny -= (float)LK_HALF;
        float pdx = 0.f, pdy = 0.f;
        // stage the 32 x 32 region that holds the 23 x 23 samples at (inx, iny): one TMA tile clamped into the image when the samples lie inside
        // it, else the reflect-101 path (origin a multiple of 4 in x)
        auto stage_j = [&](int inx, int iny) {
#ifndef VIWB_HOST_EMU
            if (tma && inx >= 0 && iny >= 0 && inx + 23 <= jc && iny + 23 <= jr) { tile_j(inx, iny); lk_tma_wait(bar + 1, phaseJ); return; }
#endif
            jx0 = (inx - LK_SLACK) & ~3; jy0 = iny - 4; jx1 = jx0 + LK_JS; jy1 = jy0 + LK_JROWS;
            lk_stage(jbuf, jimg, jstr, jc, jr, jx0, jy0, LK_JROWS, LK_JS / 4, lane);
        };
        for (int j = 0; j < max_iter; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (lk_outside(inx, iny, jc, jr)) { if (level == 0) status = false; break; }
#ifndef VIWB_HOST_EMU
            if (j_pending) { lk_tma_wait(bar + 1, phaseJ); j_pending = false; }      // the region requested at the top of the level
#endif
            if (!staged || inx < jx0 || iny < jy0 || inx + 23 > jx1 || iny + 23 > jy1) { stage_j(inx, iny); staged = true; }
            const float ja = nx - inx, jb = ny - iny;
            const int w00 = cv_round_f((1.f - ja) * (1.f - jb) * (1 << 14)), w01 = cv_round_f(ja * (1.f - jb) * (1 << 14));
            const int w10 = cv_round_f((1.f - ja) * jb * (1 << 14)), w11 = (1 << 14) - w00 - w01 - w10;
            const unsigned wt = ((unsigned)w00 & 0xffffu) | ((unsigned)w01 << 16), wb = ((unsigned)w10 & 0xffffu) | ((unsigned)w11 << 16);
            const uint8_t *q0 = jbuf + (iny - jy0) * LK_JS;
            const int xo = inx - jx0;
            int sb1 = 0, sb2 = 0;      // per lane < 14 * 2^25
#pragma unroll
            for (int s = 0; s < LK_IPL; s++) {
                if (lane + LK_W * s < LK_ITEMS) {
                    int v[7];
                    lk_sample7(q0 + irow[s] * LK_JS, xo + icol[s], wt, wb, Ic[s], v);
#pragma unroll
                    for (int k = 0; k < 7; k++) { const int diff = v[k] >> 9; sb1 += diff * gx[s][k]; sb2 += diff * gy[s][k]; }
                }
            }
            const long long b1 = lk_warp_sum(sb1), b2 = lk_warp_sum(sb2);
            const float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
            const float dx = (fA12 * fb2 - fA22 * fb1) * D, dy = (fA12 * fb1 - fA11 * fb2) * D;
            nx += dx; ny += dy;
            npx = nx + (float)LK_HALF; npy = ny + (float)LK_HALF;
            if (dx * dx + dy * dy <= eps2) break;
            if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) { npx -= dx * 0.5f; npy -= dy * 0.5f; break; }
            pdx = dx; pdy = dy;
        }
#ifndef VIWB_HOST_EMU
        if (j_pending) { lk_tma_wait(bar + 1, phaseJ); j_pending = false; }      // (no iteration ran: the requested region is still to be collected)
#endif
        if (status && level == 0) {   // L1 error of the final patch (flags without OPTFLOW_LK_GET_MIN_EIGENVALS)
            const float ex = npx - (float)LK_HALF, ey = npy - (float)LK_HALF;
            const int inx = (int)floorf(ex), iny = (int)floorf(ey);
            if (lk_outside(inx, iny, jc, jr)) { status = false; }
            else {
                if (!staged || inx < jx0 || iny < jy0 || inx + 23 > jx1 || iny + 23 > jy1) stage_j(inx, iny);
                const float ja = ex - inx, jb = ey - iny;
                const int w00 = cv_round_f((1.f - ja) * (1.f - jb) * (1 << 14)), w01 = cv_round_f(ja * (1.f - jb) * (1 << 14));
                const int w10 = cv_round_f((1.f - ja) * jb * (1 << 14)), w11 = (1 << 14) - w00 - w01 - w10;
                const unsigned wt = ((unsigned)w00 & 0xffffu) | ((unsigned)w01 << 16), wb = ((unsigned)w10 & 0xffffu) | ((unsigned)w11 << 16);
                const uint8_t *q0 = jbuf + (iny - jy0) * LK_JS;
                const int xo = inx - jx0;
                int sev = 0;
#pragma unroll
                for (int s = 0; s < LK_IPL; s++) {
                    if (lane + LK_W * s < LK_ITEMS) {
                        int v[7];
                        lk_sample7(q0 + irow[s] * LK_JS, xo + icol[s], wt, wb, Ic[s], v);
#pragma unroll
                        for (int k = 0; k < 7; k++) { const int diff = v[k] >> 9; sev += diff < 0 ? -diff : diff; }
                    }
                }
                errv = (float)lk_warp_sum(sev) * (1.f / (32 * LK_WIN * LK_WIN));
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
// maps: the level tensor maps in GLOBAL memory (written by the host before the launch), or nullptr = no TMA staging
__global__ void __launch_bounds__(32 * LK_PPB, LK_MINB) lk_track_tasks_kernel(const LkArgs *t, const LkMaps *maps) {
    extern __shared__ __align__(128) unsigned char lk_smem[];
    const int warp = threadIdx.x >> 5;
    lk_track_warp(t[blockIdx.y], maps, blockIdx.x * LK_PPB + warp, threadIdx.x & 31, lk_smem + (size_t)warp * LK_WARP_SMEM);
}
#endif

}  // namespace viwb

// viwb.cu -- host side of libviwb.so: the C ABI of include/viwb.h over the CUDA kernels in this directory.
//
// Lowers viwb_problem tables to the device layout of layout.cuh (sorting visual factors by landmark, building the
// per-frame-pair gather lists, compacting the active tangent columns), runs the fixed launch sequence of one
// Estimator::optimization() (estimator.cpp:1383-1896) for a whole batch of windows and brings results back.
// There is no host arithmetic on the data path and no CPU fallback: without a CUDA device every entry point fails.
//
// The same file compiles with -DVIWB_HOST_EMU (g++, no CUDA) into the test-only kernel-logic emulation used by
// the `not gpu` tests (tests/emu); that build is never part of libviwb.so.
#include "../../include/viwb.h"
#include "kernels_marg.cuh"
#include "kernels_fused.cuh"
#include "kernels_lk.cuh"
#include "kernels_preint.cuh"
#include "kernels_feat.cuh"
#include "kernels_init.cuh"
#include "kernels_detect.cuh"
#include "kernels_track.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

using namespace viwb;

// ====================================================================================== device abstraction
#ifdef VIWB_HOST_EMU
typedef void *stream_t;
// poison fresh 'device' memory with NaN bytes so that reads of never-written data show up in the CPU tests
static int dev_malloc(void **p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xFF, n ? n : 1); return *p ? 0 : 1; }
static void dev_free(void *p) { free(p); }
static int dev_h2d(void *d, const void *h, size_t n, stream_t) { if (n) memcpy(d, h, n); return 0; }
static int dev_d2h(void *h, const void *d, size_t n, stream_t) { if (n) memcpy(h, d, n); return 0; }
static int dev_d2d(void *d, const void *s, size_t n, stream_t) { if (n) memcpy(d, s, n); return 0; }
static int dev_sync(stream_t) { return 0; }
static int host_alloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
static void host_free(void *p) { free(p); }
static const char *dev_errstr(int) { return "emulation"; }
#define VIWB_EMU_NT 1
template <typename F>
static void emu_launch(F f, const BatchDev &bd, int gx, int gy, size_t smem_bytes, int mode) {
#ifdef VIWB_EMU_STRICT
    std::vector<double> sm((smem_bytes + 7) / 8);       // sanitizer build: exactly the dynamic shared memory the launch asks for
#else
    std::vector<double> sm(smem_bytes / 8 + 64);
#endif
    for (int by = 0; by < gy; by++) for (int bx = 0; bx < gx; bx++) f(bd, bx, by, 0, 1, sm.data(), mode);
}
#define LAUNCH(name, bd, gx, gy, nt, smem_bytes, mode, stream) emu_launch(name##_block, bd, gx, gy, smem_bytes, mode)
#define NT(n) 1
#else
#include <cuda_runtime.h>
typedef cudaStream_t stream_t;
static int dev_malloc(void **p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
static void dev_free(void *p) { cudaFree(p); }
static int dev_h2d(void *d, const void *h, size_t n, stream_t s) { return n ? (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) : 0; }
static int dev_d2h(void *h, const void *d, size_t n, stream_t s) { return n ? (int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) : 0; }
static int dev_d2d(void *d, const void *s_, size_t n, stream_t s) { return n ? (int)cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s) : 0; }
static int dev_sync(stream_t s) { return (int)cudaStreamSynchronize(s); }
static int host_alloc(void **p, size_t n) { return (int)cudaHostAlloc(p, n ? n : 1, cudaHostAllocDefault); }
static void host_free(void *p) { cudaFreeHost(p); }
static const char *dev_errstr(int e) { return cudaGetErrorString((cudaError_t)e); }
// optional per-kernel CUDA-event profiler (bench.py's live roofline timing); off by default
struct Profiler {
    bool on = false;
    std::vector<std::string> names; std::vector<double> ms; std::vector<long long> count;
    struct Rec { int idx; cudaEvent_t a, b; };
    std::vector<Rec> pending;
    int index(const char *n) { for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return (int)i; names.push_back(n); ms.push_back(0); count.push_back(0); return (int)names.size() - 1; }
    void begin(const char *n, cudaStream_t s) { if (!on) return; Rec r; r.idx = index(n); cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, s); pending.push_back(r); }
    void end(cudaStream_t s) { if (!on) return; cudaEventRecord(pending.back().b, s); }
    void collect() { for (auto &r : pending) { cudaEventSynchronize(r.b); float t = 0; cudaEventElapsedTime(&t, r.a, r.b); ms[r.idx] += t; count[r.idx]++; cudaEventDestroy(r.a); cudaEventDestroy(r.b); } pending.clear(); }
};
static Profiler g_prof;
#define DEF_KERNEL2(name, maxnt, minb) \
    __global__ void __launch_bounds__(maxnt, minb) name##_kernel(BatchDev bd, int mode) { \
        extern __shared__ double viwb_smem[]; \
        name##_block(bd, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x, viwb_smem, mode); }
#define DEF_KERNEL(name, maxnt) DEF_KERNEL2(name, maxnt, 1)
#ifndef LIN_VIS_MINB
#define LIN_VIS_MINB 3
#endif
DEF_KERNEL(vis_expand, 128)
DEF_KERNEL(setup, 128)
DEF_KERNEL(prior_setup, 256)
DEF_KERNEL2(lin_vis, 128, LIN_VIS_MINB)
#ifndef LM_MINB
#define LM_MINB 8
#endif
DEF_KERNEL2(lm_reduce, 128, LM_MINB)
DEF_KERNEL(lm_reduce_wide, 128)
#ifndef LSM_MINB
#define LSM_MINB 4      // 128 registers, 4 blocks / SM: the spills of the IMU evaluation (10 threads of the block) cost less than the idle SM (r02n: 0.38 -> 0.25 ms per launch)
#endif
DEF_KERNEL2(lin_small, 128, LSM_MINB)
DEF_KERNEL(asm_items, 128)
DEF_KERNEL(asm_items_split, 128)
#ifndef SYRK_MINB
#define SYRK_MINB 4
#endif
DEF_KERNEL2(syrk, 256, SYRK_MINB)
#ifndef SOLVE_NT
#define SOLVE_NT 384
#endif
DEF_KERNEL2(solve, SOLVE_NT, 2)
#ifndef LVL_MINB
#define LVL_MINB 4
#endif
DEF_KERNEL2(lin_vis_lm, LMB_FACTORS, LVL_MINB)
DEF_KERNEL2(lin_vis_lm_wide, LMB_FACTORS, 3)
DEF_KERNEL(asm_pairs, 128)
DEF_KERNEL(asm_pairs_wide, 128)
#ifndef PAIR_RED_NT
#define PAIR_RED_NT 256
#endif
DEF_KERNEL(pair_reduce, PAIR_RED_NT)
DEF_KERNEL2(pair_win, 256, 2)
DEF_KERNEL2(syrk_mma, SYRK_NT, 4)
DEF_KERNEL(reanchor, 32)
DEF_KERNEL2(marg_prep, 256, 2)
DEF_KERNEL2(marg_eig, 256, 3)
#ifndef MARG_TRI_NT
#define MARG_TRI_NT 256
#endif
DEF_KERNEL2(marg_tri, MARG_TRI_NT, 3)
DEF_KERNEL(marg_ql, 32)
DEF_KERNEL2(marg_apply, 128, 4)
DEF_KERNEL(outlier, 128)
#define LAUNCH(name, bd, gx, gy, nt, smem_bytes, mode, stream) \
    do { if ((gx) > 0 && (gy) > 0) { g_prof.begin((mode) == 1 ? #name "_marg" : #name, stream); name##_kernel<<<dim3((gx), (gy)), (nt), (smem_bytes), (stream)>>>(bd, mode); g_prof.end(stream); } } while (0)
#define NT(n) (n)
#endif

// one device slab + one pinned host staging slab; cached in the context so that the host-buffer entry points do not
// pay cudaMalloc / cudaHostAlloc on every call
struct Arena { char *dev = nullptr; size_t dev_cap = 0; char *host = nullptr; size_t host_cap = 0; bool busy = false; };
static void arena_release(Arena &a) { if (a.dev) dev_free(a.dev); if (a.host) host_free(a.host); a = Arena(); }
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool g_timing = getenv("VIWB_TIMING") != nullptr;
static bool g_marg_one_kernel = getenv("VIWB_MARG_ONE_KERNEL") != nullptr;      // measurement aid: tred2 + tql2 in one block-per-window kernel (marg_eig)
static bool g_no_pair_win = getenv("VIWB_PAIR_WIN") == nullptr;      // pair_win (chunk products in shared memory, one block per window) is opt-in: at the bench window's 109 KB per block it measured 0.66 ms against 0.21 + 0.14 for asm_pairs + pair_reduce (profiles/r02zd_probe.txt)
static bool g_syrk_dfma = getenv("VIWB_SYRK_DFMA") != nullptr;      // measurement aid: the 4x4 register-tiled DFMA SYRK instead of the DMMA one (profiles/: both builds of the Schur GEMM)
struct viwb_context {
    int device;
    stream_t stream;        // the stream all work of this context is issued on (own_stream unless viwb_set_stream gave another)
    stream_t own_stream;    // created by viwb_create, destroyed by viwb_destroy; a caller's stream is never destroyed here
    long long launches;
    long long h2d_bytes;    // bytes of window tables uploaded so far (the wire format, not the caller's tables)
    std::string err;
    bool attrs_set;
    Arena arena;
    struct viwb_lk_batch *lk1;      // one-stream LK state behind viwb_lk_track / viwb_track_checked
    struct viwb_detector *det1;     // one-stream detector behind viwb_set_mask / viwb_good_features_to_track
};

static int fail(viwb_context *ctx, int code, const std::string &msg) { if (ctx) ctx->err = msg; return code; }
// CUDA's current device is per host thread: every entry point binds the context's device before touching it, so that
// contexts can be driven from any thread (one context per thread at a time)
static inline void bind_device(const viwb_context *ctx) {
#ifndef VIWB_HOST_EMU
    if (ctx) { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != ctx->device) cudaSetDevice(ctx->device); }   // the host application may have switched devices in between
#else
    (void)ctx;
#endif
}
#define CK(call) do { int e_ = (call); if (e_) return fail(ctx, VIWB_ERR_CUDA, std::string(#call) + ": " + dev_errstr(e_)); } while (0)

#include "lk_host.inl"
#include "detect_host.inl"
#include "track_host.inl"

// ====================================================================================== batch
struct HostPrior { int valid, n, nb; int block_id[NB], block_idx[NB]; std::vector<double> x0, J, r; };
struct viwb_batch {
    int B;
    BatchDev bd;
    Arena arena; bool arena_cached;      // cached: borrowed from the context (host-buffer calls), else owned
    size_t out_bytes, out_off;           // staging of the results inside the pinned slab: after the inputs (the upload is not waited for, so the regions must not overlap)
    std::vector<int> prior_n;            // n of the prior each window will produce (known from the plan)
    int prior_nmax;
    WinWork *work_init_dev;              // pristine solver states, copied to bd.work at the start of every run
    std::vector<WinMeta> meta;
    std::vector<int> out_mode;        // 0: prior computed on the device, 1: input prior passes through, 2: invalid / none
    std::vector<HostPrior> in_prior;
    std::vector<int> state_sizes;
    size_t total_state;
    double algorithmic_bytes;
    bool any_marg;
    int max_iter;
    double max_time;       // max_solver_time_in_seconds (0 = off): checked on the host between rounds, like Ceres' wall-clock test
    size_t nrec_imu, nrec_wheel, nrec_plane;
};

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static void opts_from(const viwb_options *o, Opts &d) {
    d.max_num_iterations = o->max_num_iterations; d.max_invalid = o->max_num_consecutive_invalid_steps; d.jacobi_scaling = o->jacobi_scaling;
    d.function_tolerance = o->function_tolerance; d.gradient_tolerance = o->gradient_tolerance; d.parameter_tolerance = o->parameter_tolerance;
    d.initial_radius = o->initial_trust_region_radius; d.max_radius = o->max_trust_region_radius; d.min_radius = o->min_trust_region_radius;
    d.min_relative_decrease = o->min_relative_decrease; d.min_lm_diagonal = o->min_lm_diagonal; d.max_lm_diagonal = o->max_lm_diagonal;
}

extern "C" void viwb_default_options(viwb_options *o) {
    o->max_num_iterations = 8; o->max_solver_time_in_seconds = 0;
    o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
    o->max_num_consecutive_invalid_steps = 5; o->jacobi_scaling = 1;
}
extern "C" void viwb_default_globals(viwb_globals *g) {
    g->G[0] = 0; g->G[1] = 0; g->G[2] = 9.81007;
    g->vis_sqrt_info[0] = 460.0 / 1.5; g->vis_sqrt_info[1] = 0; g->vis_sqrt_info[2] = 0; g->vis_sqrt_info[3] = 460.0 / 1.5;
    g->plane_sqrt_info[0] = 100.0; g->plane_sqrt_info[1] = 100.0; g->plane_sqrt_info[2] = 20.0;
    g->huber_delta = 1.0;
}

static double window_algorithmic_bytes(const viwb_problem &p, int iters) {
    // B_solve = iters * B_iter + B_marg
    double R = 0, pg = p.num_landmarks;
    for (int b = 0; b < NB; b++) if (p.block_flags[b] & VIWB_BLOCK_PRESENT) { pg += blk_size(b); if (!(p.block_flags[b] & VIWB_BLOCK_CONSTANT)) R += blk_tsize(b); }
    const double n = (p.prior && p.prior->valid) ? p.prior->n : 0;
    const double b_iter = 112.0 * p.num_vis + 2296.0 * p.num_imu + 624.0 * p.num_wheel + 4.0 * p.num_plane + 8 * (n * n + 2 * n) + 8 * pg + 2 * 8 * (R * R + R) + 8 * pg;
    int nv0 = 0; std::vector<char> l0(p.num_landmarks > 0 ? p.num_landmarks : 1, 0);
    for (int i = 0; i < p.num_vis; i++) if (p.vis_frame_i[i] == 0) { nv0++; l0[p.vis_landmark[i]] = 1; }
    double m = 15; for (int k = 0; k < p.num_landmarks; k++) m += l0[k];
    const double nn = std::max(n, 76.0);
    const double b_marg = 112.0 * nv0 + 2296 + (p.num_wheel ? 624 : 0) + (p.num_plane ? 4 : 0) + 8 * (n * n + 2 * n) + 8 * ((m + nn) * (m + nn) + (m + nn)) + 8 * (nn * nn + nn);
    return iters * b_iter + b_marg;
}

static void batch_free(viwb_context *ctx, viwb_batch *b);

// ---- per-window lowering, phase 1: validation, sizes, active columns, marginalisation plan (no offsets yet)
struct WinLow {
    int err; bool grouped, has_prior;
    int nlist_s, nitems_s, nph_s, nlist_m, nitems_m, nph_m;
    int prior_n_out;
    bool regular;               // fused path possible: grouped table, <= LMB_FACTORS factors per landmark, at most two factors per (landmark, observer), consecutive
    int nlmb, npitems, nxrec;   // landmark blocks of lin_vis_lm, pair chunks of asm_pairs, records (two-frame factors; + the one-frame ones when the records are wide)
    int nmpitems, nmxrec;       // the same for the marginalisation of frame 0 (factors hosted there, wide records)
    bool obs_per_factor; int nobs_i, obsi_off;       // entries of the host-side observation table (wire format): one per run of factors with an identical (pts_i, velocity_i, td_i)
};
static int chunk_count(int n) { return (n + ASM_CHUNK - 1) / ASM_CHUNK; }
// columns of a visual factor's 12-double observation record (viwb.h): pts_i (0..2) pts_j (3..5) velocity_i (6,7) velocity_j (8,9) td_i (10) td_j (11)
static inline bool same_host_side(const double *a, const double *b) {      // bitwise (the wire format must reproduce the caller's table exactly)
    auto bits = [](const double *p) { unsigned long long v; memcpy(&v, p, sizeof v); return v; };
    return ((bits(a) ^ bits(b)) | (bits(a + 1) ^ bits(b + 1)) | (bits(a + 2) ^ bits(b + 2)) | (bits(a + 6) ^ bits(b + 6)) | (bits(a + 7) ^ bits(b + 7)) | (bits(a + 10) ^ bits(b + 10))) == 0;
}
static void split_obs(const double *o, double *hi, double *hj) { hi[0] = o[0]; hi[1] = o[1]; hi[2] = o[2]; hi[3] = o[6]; hi[4] = o[7]; hi[5] = o[10]; hj[0] = o[3]; hj[1] = o[4]; hj[2] = o[5]; hj[3] = o[8]; hj[4] = o[9]; hj[5] = o[11]; }

static void lower_count(const viwb_problem &p, int mf, WinMeta &m, WinLow &lo, int &out_mode) {
    memset(&m, 0, sizeof m); memset(&lo, 0, sizeof lo);
    if (p.frame_count < 0 || p.frame_count > VIWB_WINDOW_SIZE || p.num_landmarks < 0 || p.num_landmarks > VIWB_MAX_LANDMARKS) { lo.err = 1; return; }
    if (p.num_vis < 0 || p.num_imu < 0 || p.num_wheel < 0 || p.num_plane < 0) { lo.err = 5; return; }            // counts are added up as size_t below
    if ((p.num_vis && (!p.vis_type || !p.vis_landmark || !p.vis_frame_i || !p.vis_frame_j || !p.vis_obs)) || (p.num_imu && (!p.imu_frame_i || !p.imu_frame_j || !p.imu_data)) ||
        (p.num_wheel && (!p.wheel_frame_i || !p.wheel_frame_j || !p.wheel_data)) || (p.num_plane && !p.plane_frame)) { lo.err = 5; return; }
    m.nlm = p.num_landmarks; m.frame_count = p.frame_count; m.nvis = p.num_vis; m.nimu = p.num_imu; m.nwheel = p.num_wheel; m.nplane = p.num_plane;
    for (int k = 0; k < 3; k++) m.G[k] = p.globals.G[k];
    for (int k = 0; k < 4; k++) m.S_vis[k] = p.globals.vis_sqrt_info[k];
    for (int k = 0; k < 3; k++) m.w_plane[k] = p.globals.plane_sqrt_info[k];
    m.huber = p.globals.huber_delta;
    lo.has_prior = p.prior && p.prior->valid;
    if (lo.has_prior && (p.prior->n <= 0 || p.prior->n > MAXPRI || p.prior->num_blocks <= 0 || p.prior->num_blocks > NB || !p.prior->J || !p.prior->r || !p.prior->x0)) { lo.err = 2; return; }
    if (lo.has_prior) {      // every kept block once, inside [0, n), no two blocks overlapping: prior_dx and the prior's assembly trust the column map
        unsigned seen_id = 0; unsigned char used[MAXPRI]; memset(used, 0, sizeof used);
        for (int i = 0; i < p.prior->num_blocks; i++) {
            const int bq = p.prior->block_id[i], idx = p.prior->block_idx[i];
            if (bq < 0 || bq >= NB || ((seen_id >> bq) & 1u)) { lo.err = 2; return; }
            seen_id |= 1u << bq;
            const int ms = blk_msize(bq);
            if (idx < 0 || idx + ms > p.prior->n) { lo.err = 2; return; }
            for (int k = 0; k < ms; k++) { if (used[idx + k]) { lo.err = 2; return; } used[idx + k] = 1; }
        }
    }
    // visual table checks + per-list counts
    int fcnt[NFR] = {0}, pcnt[NFR * NFR] = {0}, fcnt0[NFR] = {0}, pcnt0[NFR * NFR] = {0}, ncommon0 = 0;
    bool ref[NB]; for (int k = 0; k < NB; k++) ref[k] = false;
    bool seen[NB]; for (int k = 0; k < NB; k++) seen[k] = false;
    bool any_lm0 = false;
    lo.grouped = true;
    for (int i = 0; i < p.num_vis; i++) {
        const int t = p.vis_type[i], l = p.vis_landmark[i], fi = p.vis_frame_i[i], fj = p.vis_frame_j[i];
        if (t < 0 || t > 2 || l < 0 || l >= p.num_landmarks || fi < 0 || fi > p.frame_count || fj < 0 || fj > p.frame_count) { lo.err = 3; return; }
        if (i && l < p.vis_landmark[i - 1]) lo.grouped = false;
        if (t != 2 && fi == fj) { lo.err = 3; return; }       // a two-frame factor joins two different frames (estimator.cpp:1601-1603); the pair lists have no (a, a) slot
        if (t != 2) {
            ref[fi] = ref[fj] = true; fcnt[fi]++; fcnt[fj]++;
            const int a = fi < fj ? fi : fj, c = fi < fj ? fj : fi; pcnt[a * NFR + c]++;
            if (fi == 0) { fcnt0[fi]++; fcnt0[fj]++; pcnt0[a * NFR + c]++; seen[0] = seen[fj] = true; }
        }
        ref[BLK_EX0] = true; if (t != 0) ref[BLK_EX1] = true; ref[BLK_TD] = true;
        if (fi == 0) { any_lm0 = true; ncommon0++; seen[BLK_EX0] = true; if (t != 0) seen[BLK_EX1] = true; seen[BLK_TD] = true; }
    }
    {   // every factor of a landmark is hosted by the landmark's start frame (estimator.cpp:1595-1597): lm_reduce adds the host blocks of a landmark into ONE row of W
        short host[VIWB_MAX_LANDMARKS];
        for (int k = 0; k < p.num_landmarks; k++) host[k] = -1;
        for (int i = 0; i < p.num_vis; i++) { short &h = host[p.vis_landmark[i]]; if (h < 0) h = (short)p.vis_frame_i[i]; else if (h != p.vis_frame_i[i]) { lo.err = 3; return; } }
    }
    if (lo.has_prior) for (int i = 0; i < p.prior->num_blocks; i++) ref[p.prior->block_id[i]] = true;
    for (int i = 0; i < p.num_imu; i++) { const int a = p.imu_frame_i[i], c = p.imu_frame_j[i]; if (a < 0 || a > p.frame_count || c < 0 || c > p.frame_count) { lo.err = 4; return; } ref[a] = ref[BLK_SB0 + a] = ref[c] = ref[BLK_SB0 + c] = true; }
    for (int i = 0; i < p.num_wheel; i++) { const int a = p.wheel_frame_i[i], c = p.wheel_frame_j[i]; if (a < 0 || a > p.frame_count || c < 0 || c > p.frame_count) { lo.err = 4; return; } ref[a] = ref[c] = ref[BLK_EXW] = ref[BLK_SX] = ref[BLK_SY] = ref[BLK_SW] = ref[BLK_TDW] = true; }
    for (int i = 0; i < p.num_plane; i++) { const int a = p.plane_frame[i]; if (a < 0 || a > p.frame_count) { lo.err = 4; return; } ref[a] = ref[BLK_EXW] = ref[BLK_PR] = ref[BLK_PZ] = true; }
    // active blocks (Program::RemoveFixedBlocks) and compact columns
    // Column order = elimination order of the dense Cholesky.  The speed-bias blocks come first, newest frame first: each couples
    // only with its neighbour in the IMU chain and two poses, so the factor keeps a narrow profile there (a "skyline": row i is
    // stored from its first structurally non-zero column efirst[i]; Cholesky never fills outside that envelope).
    int col = 0, amb = 0;
    bool active[NB];
    for (int k = 0; k < NB; k++) {
        m.flags[k] = p.block_flags[k] & 3u; m.mask[k] = p.subset_mask[k];
        active[k] = (p.block_flags[k] & VIWB_BLOCK_PRESENT) && !(p.block_flags[k] & VIWB_BLOCK_CONSTANT) && ref[k];
        m.tcol[k] = -1;
    }
    for (int f = VIWB_WINDOW_SIZE; f >= 0; f--) { const int k = BLK_SB0 + f; if (active[k]) { m.tcol[k] = (short)col; col += blk_tsize(k); amb += blk_size(k); } }
    for (int k = 0; k < NB; k++) if (active[k] && m.tcol[k] < 0) { m.tcol[k] = (short)col; col += blk_tsize(k); amb += blk_size(k); }
    m.nf = col; m.namb = amb + p.num_landmarks;
    {   // symbolic envelope: every factor makes the blocks it touches a clique
        int first_blk[NB];
        for (int k = 0; k < NB; k++) first_blk[k] = m.tcol[k];                 // the diagonal block itself
        auto clique = [&](const int *blks, int n) {
            int lo = TFIX;
            for (int q = 0; q < n; q++) if (blks[q] >= 0 && m.tcol[blks[q]] >= 0 && m.tcol[blks[q]] < lo) lo = m.tcol[blks[q]];
            for (int q = 0; q < n; q++) if (blks[q] >= 0 && m.tcol[blks[q]] >= 0 && lo < first_blk[blks[q]]) first_blk[blks[q]] = lo;
        };
        if (lo.has_prior) { int bl[NB]; for (int i = 0; i < p.prior->num_blocks; i++) bl[i] = p.prior->block_id[i]; clique(bl, p.prior->num_blocks); }
        for (int i = 0; i < p.num_imu; i++) { const int a = p.imu_frame_i[i], c = p.imu_frame_j[i]; const int bl[4] = {a, BLK_SB0 + a, c, BLK_SB0 + c}; clique(bl, 4); }
        for (int i = 0; i < p.num_wheel; i++) { const int bl[7] = {p.wheel_frame_i[i], p.wheel_frame_j[i], BLK_EXW, BLK_SX, BLK_SY, BLK_SW, BLK_TDW}; clique(bl, 7); }
        for (int i = 0; i < p.num_plane; i++) { const int bl[4] = {p.plane_frame[i], BLK_EXW, BLK_PR, BLK_PZ}; clique(bl, 4); }
        if (p.num_vis > 0) {       // after the landmark elimination the visual subspace (all observed poses, ex0, ex1, td) is dense
            int bl[NFR + 3], n = 0;
            for (int f = 0; f <= VIWB_WINDOW_SIZE; f++) if (ref[f]) bl[n++] = f;
            bl[n++] = BLK_EX0; bl[n++] = BLK_EX1; bl[n++] = BLK_TD;
            clique(bl, n);
        }
        int es = 0;
        for (int k = 0; k < NB; k++) if (m.tcol[k] >= 0) for (int r = 0; r < blk_tsize(k); r++) { const int i = m.tcol[k] + r; m.efirst[i] = (short)first_blk[k]; es += i - first_blk[k] + 1; }
        m.esize = es;
    }
    m.has_common = (m.tcol[BLK_EX0] >= 0 || m.tcol[BLK_EX1] >= 0 || m.tcol[BLK_TD] >= 0) ? 1 : 0;
    // marginalisation plan (estimator.cpp:1666-1893)
    m.margin_flag = -1; out_mode = 2; lo.prior_n_out = 0;
    if (mf >= 0 && p.frame_count == VIWB_WINDOW_SIZE) {
        if (mf == VIWB_MARGIN_OLD) {
            if (lo.has_prior) for (int i = 0; i < p.prior->num_blocks; i++) seen[p.prior->block_id[i]] = true;
            for (int i = 0; i < p.num_imu; i++) if (p.imu_frame_i[i] == 0 && p.imu_frame_j[i] == 1) seen[0] = seen[BLK_SB0] = seen[1] = seen[BLK_SB0 + 1] = true;
            for (int i = 0; i < p.num_wheel; i++) if (p.wheel_frame_i[i] == 0 && p.wheel_frame_j[i] == 1) seen[0] = seen[1] = seen[BLK_EXW] = seen[BLK_SX] = seen[BLK_SY] = seen[BLK_SW] = seen[BLK_TDW] = true;
            for (int i = 0; i < p.num_plane; i++) if (p.plane_frame[i] == 0) seen[0] = seen[BLK_EXW] = seen[BLK_PR] = seen[BLK_PZ] = true;
            if (seen[0] || seen[BLK_SB0] || any_lm0) { m.margin_flag = 0; out_mode = 0; }
        } else {
            for (int k = 0; k < NB; k++) seen[k] = false;
            bool has9 = false;
            if (lo.has_prior) for (int i = 0; i < p.prior->num_blocks; i++) { seen[p.prior->block_id[i]] = true; if (p.prior->block_id[i] == VIWB_WINDOW_SIZE - 1) has9 = true; }
            if (has9) { m.margin_flag = 1; out_mode = 0; } else out_mode = lo.has_prior ? 1 : 2;
        }
        if (m.margin_flag >= 0) {
            int nn = 0;
            for (int k = 0; k < NB; k++) if (seen[k]) { m.flags[k] |= 4u; const bool drop = m.margin_flag == 0 ? (k == 0 || k == BLK_SB0) : (k == VIWB_WINDOW_SIZE - 1); if (!drop) nn += blk_msize(k); }
            lo.prior_n_out = nn;
        }
    }
    // assembly plan sizes
    int nframe_ent = 0, npair_ent = 0;
    for (int a = 0; a < NFR; a++) { nframe_ent += fcnt[a]; lo.nitems_s += chunk_count(fcnt[a]); lo.nph_s = std::max(lo.nph_s, chunk_count(fcnt[a])); }
    for (int a = 0; a < NFR * NFR; a++) { npair_ent += pcnt[a]; lo.nitems_s += chunk_count(pcnt[a]); lo.nph_s = std::max(lo.nph_s, chunk_count(pcnt[a])); }
    lo.nlist_s = nframe_ent + npair_ent;
    if (m.has_common) { lo.nlist_s += p.num_vis; lo.nitems_s += chunk_count(p.num_vis); lo.nph_s = std::max(lo.nph_s, chunk_count(p.num_vis)); }
    if (m.margin_flag == 0) {
        for (int a = 0; a < NFR; a++) { lo.nlist_m += fcnt0[a]; lo.nitems_m += chunk_count(fcnt0[a]); lo.nph_m = std::max(lo.nph_m, chunk_count(fcnt0[a])); }
        for (int a = 0; a < NFR * NFR; a++) { lo.nlist_m += pcnt0[a]; lo.nitems_m += chunk_count(pcnt0[a]); lo.nph_m = std::max(lo.nph_m, chunk_count(pcnt0[a])); }
        lo.nlist_m += ncommon0; lo.nitems_m += chunk_count(ncommon0); lo.nph_m = std::max(lo.nph_m, chunk_count(ncommon0));
    }
    m.nitems = lo.nitems_s; m.nphases = lo.nph_s; m.nmitems = lo.nitems_m; m.nmphases = lo.nph_m;
    // fused-path plan (kernels_fused.cuh): sizes only; lower_fill builds the tables
    // wire format: runs of factors sharing one host-side observation (a grouped table: one run per landmark, estimator.cpp:1595-1597 takes
    // pts_i / velocity_i / td_i from the landmark's first observation for every factor); an ungrouped table keeps one entry per factor
    // (assumed here -- the plan never reads the observations --, verified when lower_fill copies them; a table that breaks it is rebuilt with one entry per factor)
    lo.obs_per_factor = !lo.grouped; lo.nobs_i = lo.grouped ? p.num_landmarks : p.num_vis;
    lo.regular = lo.grouped;
    lo.nxrec = 0; lo.npitems = 0; lo.nlmb = 0; lo.nmxrec = 0; lo.nmpitems = 0;
    for (int a = 0; a < NFR * NFR; a++) { lo.nxrec += pcnt[a]; lo.npitems += (pcnt[a] + PAIR_CHUNK - 1) / PAIR_CHUNK; }
    if (m.has_common) {      // wide records: the one-frame factors carry ex0 / ex1 / td columns; they form the "pairs" (host, host), one group per host frame
        int scnt[NFR] = {0};
        for (int i = 0; i < p.num_vis; i++) if (p.vis_type[i] == 2) scnt[p.vis_frame_i[i]]++;
        for (int a = 0; a < NFR; a++) { lo.nxrec += scnt[a]; lo.npitems += (scnt[a] + PAIR_CHUNK - 1) / PAIR_CHUNK; }
    }
    if (m.margin_flag == 0) {
        int two = 0;
        for (int a = 0; a < NFR * NFR; a++) { two += pcnt0[a]; lo.nmpitems += (pcnt0[a] + PAIR_CHUNK - 1) / PAIR_CHUNK; }
        lo.nmxrec = ncommon0; lo.nmpitems += (ncommon0 - two + PAIR_CHUNK - 1) / PAIR_CHUNK;
    }
    if (lo.regular) {
        int in_blk = 0, i = 0; bool open = false;
        for (int k = 0; k < p.num_landmarks && lo.regular; k++) {
            int cnt = 0, seenj[NFR], ndup[NFR];
            for (int q = 0; q < NFR; q++) { seenj[q] = -2; ndup[q] = 0; }
            for (; i < p.num_vis && p.vis_landmark[i] == k; i++, cnt++) {
                if (p.vis_type[i] == 2) continue;
                const int fj = p.vis_frame_j[i];
                if (ndup[fj] == 0) { ndup[fj] = 1; seenj[fj] = i; }
                else if (ndup[fj] == 1 && seenj[fj] == i - 1) ndup[fj] = 2;
                else lo.regular = false;
            }
            if (cnt > LMB_FACTORS) lo.regular = false;
            if (!open || in_blk + cnt > LMB_FACTORS) { lo.nlmb++; in_blk = 0; open = true; }
            in_blk += cnt;
        }
    }
}

// ---- phase 3: fill this window's slices of the (pinned) staging arrays; every offset is final
struct HostArrays {
    WinMeta *meta; PriorDev *prior; int *vis_code, *vis_lm, *vis_oi; double *obs_i, *obs_j; int *lm_win, *lm_fptr;
    AsmItem *items; int *asm_list; int *imu_fi, *imu_fj, *imu_win, *wheel_fi, *wheel_fj, *wheel_win, *plane_f, *plane_win;
    double *imu_data, *wheel_data, *prior_J, *prior_r, *prior_x0, *x_init; WinWork *work;
    int nitems_solve_total;
    int *vis_pos; LmbDesc *lmb_desc; AsmItem *pitems; int *mvis_pos; AsmItem *mpitems;
    unsigned char *obs_viol;        // [B] host only: set by lower_fill when a grouped table does not share one host-side observation per landmark
};
static void emit_lists(const int *type, const int *fi, const int *fj, int nvis, bool only_host0, int has_common, int w, AsmItem *items, int *list) {
    // counting sort of the (factor, role) entries into frame lists, pair lists and the common list, then chunking
    int fcnt[NFR] = {0}, pcnt[NFR * NFR] = {0}, ncom = 0;
    for (int i = 0; i < nvis; i++) {
        if (only_host0 && fi[i] != 0) continue;
        if (type[i] != 2) { fcnt[fi[i]]++; fcnt[fj[i]]++; const int a = fi[i] < fj[i] ? fi[i] : fj[i], c = fi[i] < fj[i] ? fj[i] : fi[i]; pcnt[a * NFR + c]++; }
        ncom++;
    }
    int foff[NFR], poff[NFR * NFR], pos = 0;
    for (int a = 0; a < NFR; a++) { foff[a] = pos; pos += fcnt[a]; }
    for (int a = 0; a < NFR * NFR; a++) { poff[a] = pos; pos += pcnt[a]; }
    const int coff = pos;
    int fpos[NFR], ppos[NFR * NFR], cpos = coff;
    memcpy(fpos, foff, sizeof fpos); memcpy(ppos, poff, sizeof ppos);
    for (int i = 0; i < nvis; i++) {
        if (only_host0 && fi[i] != 0) continue;
        if (type[i] != 2) {
            list[fpos[fi[i]]++] = (i << 1) | 0; list[fpos[fj[i]]++] = (i << 1) | 1;
            const int a = fi[i] < fj[i] ? fi[i] : fj[i], c = fi[i] < fj[i] ? fj[i] : fi[i];
            list[ppos[a * NFR + c]++] = (i << 1) | (fi[i] == a ? 0 : 1);
        }
        if (has_common) list[cpos++] = i << 1;
    }
    int ni = 0;
    auto emit = [&](int kind, int a, int c, int off, int cnt) {
        for (int c0 = 0, ph = 0; c0 < cnt; c0 += ASM_CHUNK, ph++) {
            AsmItem &it = items[ni++];
            it.kind = kind; it.win = w; it.a = a; it.b = c; it.lo = off + c0; it.hi = off + std::min(cnt, c0 + (int)ASM_CHUNK); it.phase = ph; it.has_common = has_common;
        }
    };
    for (int a = 0; a < NFR; a++) emit(ITEM_FRAME, a, a, foff[a], fcnt[a]);
    for (int a = 0; a < NFR; a++) for (int c = a + 1; c < NFR; c++) emit(ITEM_PAIR, a, c, poff[a * NFR + c], pcnt[a * NFR + c]);
    if (has_common) emit(ITEM_COMMON, 0, 0, coff, ncom);
}

static void lower_fill(const viwb_problem &p, const double *state, int w, const WinMeta &m, const WinLow &lo, const HostArrays &h, double init_radius) {
    h.meta[w] = m;
    // visual factors grouped by landmark
    // (the per-factor type / frames / duplicate code travel as ONE code word and are unpacked on the device: the tables below are host scratch)
    static thread_local std::vector<int> scratch; static thread_local std::vector<unsigned char> scratch_d;
    scratch.resize((size_t)3 * p.num_vis + 1); scratch_d.assign((size_t)p.num_vis + 1, 0);
    int *vt = scratch.data(), *vi = vt + p.num_vis, *vj = vi + p.num_vis, *vl = h.vis_lm + m.vis_off;
    unsigned char *vd = scratch_d.data();
    double *oj = h.obs_j + (size_t)m.vis_off * 6, *oi = h.obs_i + (size_t)lo.obsi_off * 6;
    int *voi = h.vis_oi + m.vis_off;
    if (lo.grouped) {
        memcpy(vt, p.vis_type, sizeof(int) * p.num_vis); memcpy(vl, p.vis_landmark, sizeof(int) * p.num_vis);
        memcpy(vi, p.vis_frame_i, sizeof(int) * p.num_vis); memcpy(vj, p.vis_frame_j, sizeof(int) * p.num_vis);
        int run = -1;
        for (int i = 0; i < p.num_vis; i++) {
            const double *o = p.vis_obs + (size_t)i * 12;
            double hi[6];
            split_obs(o, hi, oj + (size_t)i * 6);
            if (lo.obs_per_factor) { memcpy(oi + (size_t)i * 6, hi, sizeof hi); voi[i] = lo.obsi_off + i; continue; }
            if (i == 0 || vl[i] != vl[i - 1]) { run++; memcpy(oi + (size_t)run * 6, hi, sizeof hi); }
            else if (!same_host_side(o, o - 12)) h.obs_viol[w] = 1;          // two factors of one landmark with different host-side observations: the caller of lower_fill rebuilds
            voi[i] = lo.obsi_off + run;
        }
    } else {
        std::vector<int> order(p.num_vis);
        for (int i = 0; i < p.num_vis; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return p.vis_landmark[a] < p.vis_landmark[c]; });
        for (int i = 0; i < p.num_vis; i++) {
            const int sidx = order[i]; vt[i] = p.vis_type[sidx]; vl[i] = p.vis_landmark[sidx]; vi[i] = p.vis_frame_i[sidx]; vj[i] = p.vis_frame_j[sidx];
            split_obs(p.vis_obs + (size_t)sidx * 12, oi + (size_t)i * 6, oj + (size_t)i * 6); voi[i] = lo.obsi_off + i;
        }
    }
    // landmark -> factor ranges (global factor indices)
    { int *fp = h.lm_fptr + m.lm_off; int run = m.vis_off, i = 0;
      for (int k = 0; k < p.num_landmarks; k++) { fp[k] = run; while (i < p.num_vis && vl[i] == k) { i++; run++; } h.lm_win[m.lm_off + k] = w; } }
    // assembly plan
    // frame-pair order of the X records (counting sort by (host, observer); the one-frame factors of WIDE records form the "pair" (host, host)),
    // pair chunks; for the solver (all factors) and for the marginalisation of frame 0 (the factors hosted there)
    auto pair_tables = [&](bool marg, int *vp, AsmItem *pit, int base, int wide) {
        int pcnt[NFR * NFR] = {0}, poff[NFR * NFR];
        for (int i = 0; i < p.num_vis; i++) {
            if (marg && vi[i] != 0) continue;
            if (vt[i] != 2) pcnt[vi[i] * NFR + vj[i]]++; else if (wide) pcnt[vi[i] * NFR + vi[i]]++;
        }
        int pos = 0, ni = 0;
        for (int a = 0; a < NFR * NFR; a++) {
            poff[a] = pos;
            for (int c0 = 0, ph = 0; c0 < pcnt[a]; c0 += PAIR_CHUNK, ph++) {
                AsmItem &it = pit[ni++];
                it.kind = ITEM_PAIR; it.win = w; it.a = a / NFR; it.b = a % NFR; it.lo = pos + c0; it.hi = pos + std::min(pcnt[a], c0 + (int)PAIR_CHUNK); it.phase = ph; it.has_common = wide; it.base = base;
            }
            pos += pcnt[a];
        }
        for (int i = 0; i < p.num_vis; i++) {
            if (marg && vi[i] != 0) { vp[i] = -1; continue; }
            if (vt[i] != 2) vp[i] = poff[vi[i] * NFR + vj[i]]++; else if (wide) vp[i] = poff[vi[i] * NFR + vi[i]]++; else vp[i] = -1;
        }
    };
    if (m.fused) pair_tables(false, h.vis_pos + m.vis_off, h.pitems + m.pitem_off, m.xrec_off, m.has_common);
    if (m.mfused) pair_tables(true, h.mvis_pos + m.vis_off, h.mpitems + m.mpitem_off, m.mxrec_off, 1);
    if (m.fused || m.mfused) {
        // duplicate codes (two factors of one landmark observed from the same frame: left and right camera, consecutive), landmark blocks
        for (int i = 0; i < p.num_vis; i++) {
            if (vt[i] == 2) continue;
            if (i > 0 && vt[i - 1] != 2 && vl[i - 1] == vl[i] && vj[i - 1] == vj[i]) { vd[i - 1] = 1; vd[i] = 2; }
        }
        LmbDesc *ld = h.lmb_desc + m.lmb_off;
        int nb = 0, in_blk = 0, i = 0; bool open = false;
        for (int k = 0; k < p.num_landmarks; k++) {
            const int first = i;
            int cnt = 0;
            for (; i < p.num_vis && vl[i] == k; i++) cnt++;
            if (!open || in_blk + cnt > LMB_FACTORS) { LmbDesc &d = ld[nb++]; memset(&d, 0, sizeof d); d.win = w; d.k0 = m.lm_off + k; d.f0 = m.vis_off + first; in_blk = 0; open = true; }
            ld[nb - 1].k1 = m.lm_off + k + 1;           // the open block ends after this landmark
            in_blk += cnt; ld[nb - 1].nf = in_blk;
        }
    }
    { int *vc = h.vis_code + m.vis_off; for (int i = 0; i < p.num_vis; i++) vc[i] = vt[i] | (vi[i] << 2) | (vj[i] << 6) | ((int)vd[i] << 10); }
    if (!m.fused)
    emit_lists(vt, vi, vj, p.num_vis, false, m.has_common, w, h.items + m.item_off, h.asm_list + m.list_off);
    if (m.margin_flag == 0 && !m.mfused) emit_lists(vt, vi, vj, p.num_vis, true, 1, w, h.items + h.nitems_solve_total + m.mitem_off, h.asm_list + m.mlist_off);

    // small factors
    for (int i = 0; i < p.num_imu; i++) { h.imu_fi[m.imu_off + i] = p.imu_frame_i[i]; h.imu_fj[m.imu_off + i] = p.imu_frame_j[i]; h.imu_win[m.imu_off + i] = w; }
    if (p.num_imu) memcpy(h.imu_data + (size_t)m.imu_off * 287, p.imu_data, sizeof(double) * 287 * p.num_imu);
    for (int i = 0; i < p.num_wheel; i++) { h.wheel_fi[m.wheel_off + i] = p.wheel_frame_i[i]; h.wheel_fj[m.wheel_off + i] = p.wheel_frame_j[i]; h.wheel_win[m.wheel_off + i] = w; }
    if (p.num_wheel) memcpy(h.wheel_data + (size_t)m.wheel_off * 78, p.wheel_data, sizeof(double) * 78 * p.num_wheel);
    for (int i = 0; i < p.num_plane; i++) { h.plane_f[m.plane_off + i] = p.plane_frame[i]; h.plane_win[m.plane_off + i] = w; }
    // prior
    if (m.prior_idx >= 0) {
        const viwb_prior &pr = *p.prior; PriorDev &pd = h.prior[m.prior_idx];
        memcpy(h.prior_J + pd.J_off, pr.J, sizeof(double) * pr.n * pr.n); memcpy(h.prior_r + pd.r_off, pr.r, sizeof(double) * pr.n);
        memcpy(h.prior_x0 + pd.x0_off, pr.x0, sizeof(double) * SFIX);
    }
    memcpy(h.x_init + m.state_off, state, sizeof(double) * (SFIX + p.num_landmarks));
    WinWork &ww = h.work[w]; memset(&ww, 0, sizeof ww);
    ww.status = ST_RUNNING; ww.phase = PH_INIT; ww.first = 1; ww.radius = init_radius; ww.mu = 1e-8; ww.mu_lin = 1e-8; ww.term = VIWB_NO_CONVERGENCE;
}

template <typename F> static void parallel_for(int n, F f) {
    // host threads of the lowering: VIWB_HOST_THREADS (a caller that drives several contexts from several threads divides the cores among them), else
    // all hardware threads up to 16
    const char *env_nt = getenv("VIWB_HOST_THREADS");
    const int cap = (env_nt && atoi(env_nt) > 0) ? atoi(env_nt) : 16;
    int nt = (int)std::thread::hardware_concurrency(); if (nt > cap) nt = cap; if (nt < 1) nt = 1; if (nt > n / 8) nt = n / 8;
    if (nt <= 1) { for (int i = 0; i < n; i++) f(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=]() { for (int i = t; i < n; i += nt) f(i); });
    for (auto &t : th) t.join();
}

static int batch_build(viwb_context *ctx, int B, const viwb_problem *problems, const double *const *states,
                       const viwb_options *options, const int32_t *margin_flags, viwb_batch **out, bool use_cached = false, bool obs_per_factor = false) {
    bind_device(ctx);
    if (B <= 0 || !problems || !states) return fail(ctx, VIWB_ERR_INVALID, "empty batch");
    const double t_start = now_ms();
    viwb_batch *b = new viwb_batch();
    b->B = B; b->any_marg = false; b->algorithmic_bytes = 0; b->arena_cached = false; b->out_bytes = 0; b->prior_nmax = 0; b->work_init_dev = nullptr;
    viwb_options defopt; viwb_default_options(&defopt);
    const viwb_options *opt = options ? options : &defopt;
    b->max_iter = opt->max_num_iterations;
    b->max_time = opt->max_solver_time_in_seconds;
    BatchDev &bd = b->bd; memset(&bd, 0, sizeof bd);
    bd.B = B; opts_from(opt, bd.opt);
    b->meta.resize(B); b->out_mode.assign(B, 2); b->in_prior.resize(B); b->state_sizes.resize(B); b->prior_n.assign(B, 0);
    std::vector<WinLow> low(B);
    // ---- phase 1 (parallel): sizes and plans
    parallel_for(B, [&](int w) { lower_count(problems[w], margin_flags ? margin_flags[w] : -1, b->meta[w], low[w], b->out_mode[w]); });
    const double t_counted = now_ms();
    if (obs_per_factor) for (int w = 0; w < B; w++) { low[w].obs_per_factor = true; low[w].nobs_i = problems[w].num_vis < 0 ? 0 : problems[w].num_vis; }
    for (int w = 0; w < B; w++) if (low[w].err) { const int e = low[w].err; batch_free(ctx, b); return fail(ctx, VIWB_ERR_INVALID, e == 1 ? "bad frame_count / num_landmarks" : e == 2 ? "bad prior" : e == 3 ? "bad visual factor table" : e == 5 ? "negative factor count or missing table" : "bad factor frame index"); }
    // fused solver linearisation (kernels_fused.cuh) when every window qualifies; then the solver's gather lists are not built at all
    // (decided per window, so that a window takes the same path -- and gives the same bits -- whatever else the batch holds)
    const bool allow_fused = getenv("VIWB_NO_FUSED") == nullptr;
    int n_unfused = 0, n_fused_wide = 0, n_fused_compact = 0, n_mfused = 0, n_munfused = 0;
    for (int w = 0; w < B; w++) {
        const bool fused = allow_fused && low[w].regular && low[w].npitems > 0;      // (a window without two-frame factors has nothing to fuse)
        const bool mfused = allow_fused && low[w].regular && b->meta[w].margin_flag == 0 && low[w].nmpitems > 0;
        b->meta[w].fused = fused ? 1 : 0; b->meta[w].mfused = mfused ? 1 : 0;
        if (fused) { low[w].nitems_s = 0; low[w].nlist_s = 0; low[w].nph_s = 0; b->meta[w].nitems = 0; b->meta[w].nphases = 0; if (b->meta[w].has_common) n_fused_wide++; else n_fused_compact++; }
        else { low[w].npitems = 0; low[w].nxrec = 0; n_unfused++; }
        if (mfused) { low[w].nitems_m = 0; low[w].nlist_m = 0; low[w].nph_m = 0; b->meta[w].nmitems = 0; b->meta[w].nmphases = 0; n_mfused++; }
        else { low[w].nmpitems = 0; low[w].nmxrec = 0; if (b->meta[w].margin_flag == 0) n_munfused++; }
        if (!fused && !mfused) low[w].nlmb = 0;
    }
    // ---- phase 2: offsets
    size_t nlmb = 0, npit = 0, nxr = 0, nmpit = 0, nmxr = 0;      // nxr / nmxr: record regions in DOUBLES (solver / marginalisation share one buffer)
    size_t nobsi = 0;
    size_t nstate = 0, nvis = 0, nlm = 0, nimu = 0, nwheel = 0, nplane = 0, nlist = 0, nit_s = 0, nit_m = 0, npri = 0, npJ = 0, npr = 0;
    for (int w = 0; w < B; w++) {
        const viwb_problem &p = problems[w]; WinMeta &m = b->meta[w]; const WinLow &lo = low[w];
        m.state_off = (int)nstate; nstate += SFIX + p.num_landmarks; b->state_sizes[w] = SFIX + p.num_landmarks;
        m.vis_off = (int)nvis; nvis += p.num_vis; m.lm_off = (int)nlm; nlm += p.num_landmarks;
        low[w].obsi_off = (int)nobsi; nobsi += lo.nobs_i;
        m.imu_off = (int)nimu; nimu += p.num_imu; m.wheel_off = (int)nwheel; nwheel += p.num_wheel; m.plane_off = (int)nplane; nplane += p.num_plane;
        m.item_off = (int)nit_s; nit_s += lo.nitems_s; m.list_off = (int)nlist; nlist += lo.nlist_s;
        m.prior_idx = lo.has_prior ? (int)npri++ : -1;
        m.lmb_off = (int)nlmb; m.nlmb = lo.nlmb; nlmb += lo.nlmb; m.pitem_off = (int)npit; m.npitems = lo.npitems; npit += lo.npitems; m.xrec_off = (int)nxr; m.nxrec = lo.nxrec; nxr += (size_t)lo.nxrec * (m.has_common ? (int)XL<true>::REC : (int)XL<false>::REC);
        m.mpitem_off = (int)nmpit; m.nmpitems = lo.nmpitems; nmpit += lo.nmpitems; m.mxrec_off = (int)nmxr; m.nmxrec = lo.nmxrec; nmxr += (size_t)lo.nmxrec * XL<true>::REC;
        b->prior_n[w] = lo.prior_n_out; if (m.margin_flag >= 0) b->any_marg = true;
        b->prior_nmax = std::max(b->prior_nmax, lo.prior_n_out);
        b->algorithmic_bytes += window_algorithmic_bytes(p, opt->max_num_iterations);
    }
    for (int w = 0; w < B; w++) { WinMeta &m = b->meta[w]; const WinLow &lo = low[w]; m.mitem_off = (int)nit_m; nit_m += lo.nitems_m; m.mlist_off = (int)nlist; nlist += lo.nlist_m; }
    if (nxr > 0x7fffffff || nmxr > 0x7fffffff || nstate > 0x7fffffff || nvis * 12 > 0x7fffffffull * 4 || nlist > 0x7fffffff) { batch_free(ctx, b); return fail(ctx, VIWB_ERR_INVALID, "batch too large"); }
    std::vector<PriorDev> priors(npri);
    for (int w = 0; w < B; w++) if (b->meta[w].prior_idx >= 0) {
        const viwb_prior &pr = *problems[w].prior; PriorDev &pd = priors[b->meta[w].prior_idx]; memset(&pd, 0, sizeof pd);
        pd.n = pr.n; pd.nb = pr.num_blocks;
        for (int i = 0; i < pr.num_blocks; i++) { pd.block_id[i] = pr.block_id[i]; pd.block_idx[i] = pr.block_idx[i]; }
        pd.J_off = (int)npJ; npJ += (size_t)pr.n * pr.n; pd.r_off = (int)npr; npr += pr.n; pd.x0_off = (int)(SFIX * (size_t)b->meta[w].prior_idx);
        HostPrior &hp = b->in_prior[w]; hp.valid = 1; hp.n = pr.n; hp.nb = pr.num_blocks;
        memcpy(hp.block_id, pd.block_id, sizeof hp.block_id); memcpy(hp.block_idx, pd.block_idx, sizeof hp.block_idx);
        if (b->out_mode[w] == 1) { hp.x0.assign(pr.x0, pr.x0 + SFIX); hp.J.assign(pr.J, pr.J + (size_t)pr.n * pr.n); hp.r.assign(pr.r, pr.r + pr.n); }
    }
    bd.nvis_total = (int)nvis; bd.nlm_total = (int)nlm; bd.nimu_total = (int)nimu; bd.nwheel_total = (int)nwheel; bd.nplane_total = (int)nplane; bd.nprior = (int)npri;
    bd.nitems_solve = (int)nit_s; bd.nitems_marg = (int)nit_m;
    bd.rec_stride_solve = VREC_COMPACT;
    bd.n_unfused = n_unfused; bd.nlmb_total = (int)nlmb; bd.npitems_total = (int)npit; bd.nmpitems_total = (int)nmpit;
    bd.pitems_max = 1; for (int w = 0; w < B; w++) bd.pitems_max = std::max(bd.pitems_max, std::max(b->meta[w].npitems, b->meta[w].nmpitems));
    {   // pair_win keeps a window's chunk products in shared memory: possible when the largest window fits next to a second resident block
        size_t need_s = 0, need_m = 0;
        for (int w = 0; w < B; w++) {
            const WinMeta &mw = b->meta[w];
            if (mw.fused) need_s = std::max(need_s, pair_win_smem_bytes(mw.npitems, mw.has_common ? (int)XL<true>::OUT : (int)XL<false>::OUT));
            if (mw.mfused) need_m = std::max(need_m, pair_win_smem_bytes(mw.nmpitems, (int)XL<true>::OUT));
        }
        const size_t cap = 110 * 1024;
        bd.pwin_smem = (!g_no_pair_win && need_s > 0 && need_s <= cap) ? (int)need_s : 0;
        bd.pwin_smem_marg = (!g_no_pair_win && need_m > 0 && need_m <= cap) ? (int)need_m : 0;
    }
    bd.n_fused_wide = n_fused_wide; bd.n_fused_compact = n_fused_compact; bd.n_mfused = n_mfused; bd.n_munfused = n_munfused;
    bd.pout_stride = n_fused_wide ? (int)XL<true>::OUT : (int)XL<false>::OUT;
    bd.marg_nmax = b->prior_nmax;
    bd.env_max = 1; for (int w = 0; w < B; w++) bd.env_max = std::max(bd.env_max, b->meta[w].esize);
    for (int w = 0; w < B; w++) if (b->meta[w].has_common) bd.rec_stride_solve = VREC;
    b->total_state = nstate;
    // ---- placement (inputs first, then work arrays)
    struct Ent { void **field; size_t bytes; bool input; size_t off; };
    std::vector<Ent> ents;
    auto IN = [&](auto **field, size_t count) { ents.push_back({(void **)field, count * sizeof(**field), true, 0}); };
    auto WK = [&](auto **field, size_t count) { ents.push_back({(void **)field, (count ? count : 1) * sizeof(**field), false, 0}); };
    IN(&bd.meta, B); IN(&bd.prior, npri); IN(&bd.vis_code, nvis); IN(&bd.vis_lm, nvis); IN(&bd.vis_oi, nvis);
    IN(&bd.obs_i, nobsi * 6); IN(&bd.obs_j, nvis * 6); IN(&bd.lm_win, nlm); IN(&bd.lm_fptr, nlm + 1); IN(&bd.items, nit_s + nit_m); IN(&bd.asm_list, nlist);
    IN(&bd.imu_fi, nimu); IN(&bd.imu_fj, nimu); IN(&bd.imu_win, nimu); IN(&bd.wheel_fi, nwheel); IN(&bd.wheel_fj, nwheel); IN(&bd.wheel_win, nwheel);
    IN(&bd.plane_f, nplane); IN(&bd.plane_win, nplane); IN(&bd.imu_data, nimu * 287); IN(&bd.wheel_data, nwheel * 78);
    IN(&bd.prior_J, npJ); IN(&bd.prior_r, npr); IN(&bd.prior_x0, npri * SFIX); IN(&bd.x_init, nstate); IN(&b->work_init_dev, B);
    IN(&bd.vis_pos, npit ? nvis : 0); IN(&bd.lmb_desc, nlmb); IN(&bd.pitems, npit);
    IN(&bd.mvis_pos, nmpit ? nvis : 0); IN(&bd.mpitems, nmpit);
    const size_t nvec = (size_t)B * TFIX + nlm;
    WK(&bd.work, B); WK(&bd.x_cur, nstate); WK(&bd.x_cand, nstate); WK(&bd.x_before, nstate);
    WK(&bd.vis_type, nvis); WK(&bd.vis_fi, nvis); WK(&bd.vis_fj, nvis); WK(&bd.vis_win, nvis); WK(&bd.vis_dup, nvis); WK(&bd.vis_obs, nvis * 12);
    WK(&bd.vis_rec, nvis * VREC); WK(&bd.vis_cost, nvis);
    WK(&bd.xrec, std::max(nxr, nmxr)); WK(&bd.pair_out, npit * (size_t)bd.pout_stride); WK(&bd.mpair_out, nmpit * (size_t)XL<true>::OUT); WK(&bd.pair_red, (npit || nmpit) ? (size_t)B * PAIR_RED : 0);
    WK(&bd.lm_a, nlm); WK(&bd.lm_g, nlm); WK(&bd.lm_gamma, nlm); WK(&bd.lm_scale, nlm); WK(&bd.lm_cost, nlm); WK(&bd.lm_W, nlm * VSUB); WK(&bd.lm_outlier, nlm);
    WK(&bd.imu_S, nimu * 225); WK(&bd.wheel_S, nwheel * 36); WK(&bd.imu_rec, nimu * IMU_REC); WK(&bd.wheel_rec, nwheel * WHEEL_REC); WK(&bd.plane_rec, nplane * PLANE_REC);
    WK(&bd.prior_A, npJ); WK(&bd.prior_res, npr); WK(&bd.prior_g, npr);
    WK(&bd.Hpk, (size_t)B * (TFIX * (TFIX + 1) / 2)); WK(&bd.gpk, (size_t)B * TFIX); WK(&bd.gfix, (size_t)B * (TFIX + 8));
    WK(&bd.asm_out, (nit_s + nit_m) * ASM_STRIDE); WK(&bd.Tvis, (size_t)B * VSUB * VSUB); WK(&bd.tvec, (size_t)B * VSUB);
    WK(&bd.v_scale, nvec); WK(&bd.v_D, nvec); WK(&bd.v_sgrad, nvec); WK(&bd.v_gn, nvec); WK(&bd.v_wug, nlm); WK(&bd.v_wun, nlm);
    // marginalisation outputs / scratch: sized by the largest prior this batch produces, and not at all when no window marginalises
    const size_t mJ = b->any_marg ? (size_t)b->prior_nmax * b->prior_nmax : 0, mB = b->any_marg ? (size_t)B : 0;
    WK(&bd.marg_J, mB * mJ); WK(&bd.marg_r, mB * MAXPRI); WK(&bd.marg_x0, mB * SFIX);
    WK(&bd.marg_hdr, mB * (3 + 2 * NB)); WK(&bd.marg_de, mB * 2 * MAXPRI); WK(&bd.marg_rot, mB * 2 * marg_rot_cap(b->prior_nmax)); WK(&bd.marg_sweep, mB * (2 * marg_sweep_cap(b->prior_nmax) + 2));
    size_t tot = 0;
    for (auto &e : ents) if (e.input) { e.off = tot; tot += align_up(e.bytes); }
    const size_t in_bytes = tot;
    for (auto &e : ents) if (!e.input) { e.off = tot; tot += align_up(e.bytes); }
    b->out_bytes = align_up(nstate * 8) + align_up(sizeof(WinWork) * B) + align_up((size_t)B * (3 + 2 * NB) * 4) + align_up((size_t)B * MAXPRI * 8) +
                   align_up((size_t)B * SFIX * 8) + align_up((size_t)B * b->prior_nmax * b->prior_nmax * 8);
    b->out_off = align_up(in_bytes);
    const size_t host_need = b->out_off + b->out_bytes;
    Arena *ar;
    if (use_cached && !ctx->arena.busy) { ar = &ctx->arena; b->arena_cached = true; ctx->arena.busy = true; } else { ar = &b->arena; b->arena_cached = false; }
    if (ar->dev_cap < tot) { if (ar->dev) dev_free(ar->dev); ar->dev = nullptr; ar->dev_cap = 0; void *d = nullptr; int e = dev_malloc(&d, tot + tot / 8); if (e) { batch_free(ctx, b); return fail(ctx, VIWB_ERR_CUDA, std::string("device slab: ") + dev_errstr(e)); } ar->dev = (char *)d; ar->dev_cap = tot + tot / 8; }
    if (ar->host_cap < host_need) { if (ar->host) host_free(ar->host); ar->host = nullptr; ar->host_cap = 0; void *hm = nullptr; int e = host_alloc(&hm, host_need + host_need / 8); if (e) { batch_free(ctx, b); return fail(ctx, VIWB_ERR_CUDA, "pinned staging slab"); } ar->host = (char *)hm; ar->host_cap = host_need + host_need / 8; }
    if (b->arena_cached) b->arena = *ar;     // a view; ownership stays with the context
    const double t_alloc = now_ms();
    // host views of the input arrays inside the staging slab
    HostArrays h; memset(&h, 0, sizeof h);
    { size_t k = 0; char *hb = ar->host;
      auto HP = [&](auto *&dst) { dst = (typename std::remove_reference<decltype(dst)>::type)(hb + ents[k].off); k++; };
      HP(h.meta); HP(h.prior); HP(h.vis_code); HP(h.vis_lm); HP(h.vis_oi); HP(h.obs_i); HP(h.obs_j); HP(h.lm_win); HP(h.lm_fptr); HP(h.items); HP(h.asm_list);
      HP(h.imu_fi); HP(h.imu_fj); HP(h.imu_win); HP(h.wheel_fi); HP(h.wheel_fj); HP(h.wheel_win); HP(h.plane_f); HP(h.plane_win); HP(h.imu_data); HP(h.wheel_data);
      HP(h.prior_J); HP(h.prior_r); HP(h.prior_x0); HP(h.x_init); HP(h.work);
      HP(h.vis_pos); HP(h.lmb_desc); HP(h.pitems); HP(h.mvis_pos); HP(h.mpitems); }
    h.nitems_solve_total = (int)nit_s;
    for (auto &e : ents) *e.field = ar->dev + e.off;
    if (npri) memcpy(h.prior, priors.data(), sizeof(PriorDev) * npri);
    h.lm_fptr[nlm] = (int)nvis;
    // ---- phase 3 (parallel): fill
    std::vector<unsigned char> obs_viol(B, 0); h.obs_viol = obs_viol.data();
    parallel_for(B, [&](int w) { lower_fill(problems[w], states[w], w, b->meta[w], low[w], h, opt->initial_trust_region_radius); });
    if (!obs_per_factor && std::find(obs_viol.begin(), obs_viol.end(), 1) != obs_viol.end()) {      // rare: rebuild with one host-side entry per factor
        batch_free(ctx, b);
        return batch_build(ctx, B, problems, states, options, margin_flags, out, use_cached, true);
    }
    const double t_filled = now_ms();
    { int e = dev_h2d(ar->dev, ar->host, in_bytes, ctx->stream); if (e) { batch_free(ctx, b); return fail(ctx, VIWB_ERR_CUDA, std::string("H2D: ") + dev_errstr(e)); } }
    ctx->h2d_bytes += (long long)in_bytes;
    if (nvis) { LAUNCH(vis_expand, bd, B, 1, NT(128), 0, 0, ctx->stream); ctx->launches++; }      // code words + split observations -> the tables the kernels read
    // (no wait here: the kernels of the run are queued behind the upload on the same stream, the results are staged in their own region of the slab, and the
    //  slab is next written by the host only after a fetch has synchronised)
    if (g_timing) fprintf(stderr, "[viwb] build B=%d: plan+slabs %.2f ms (count %.2f), fill %.2f ms (%.1f MB), h2d %.2f ms\n", B, t_alloc - t_start, t_counted - t_start, t_filled - t_alloc, in_bytes / 1e6, now_ms() - t_filled);
    *out = b;
    return 0;
}

static int ensure_attrs(viwb_context *ctx) {
#ifndef VIWB_HOST_EMU
    if (!ctx->attrs_set) {
        CK(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(solve_smem_doubles(SOLVE_NT, TFIX * (TFIX + 1) / 2) * 8)));
        CK(cudaFuncSetAttribute(lin_vis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(lin_vis_smem_doubles(128, VREC) * 8)));
        CK(cudaFuncSetAttribute(lin_vis_lm_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(lin_vis_lm_smem_doubles(true) * 8)));
        CK(cudaFuncSetAttribute(syrk_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(syrk_mma_smem_doubles() * 8)));
        CK(cudaFuncSetAttribute(pair_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CK(cudaFuncSetAttribute(marg_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(marg_prep_smem_doubles(256, 100) * 8)));
        CK(cudaFuncSetAttribute(marg_tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(marg_eig_smem_doubles(256, 100) * 8)));
        CK(cudaFuncSetAttribute(marg_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(marg_apply_smem_doubles(128, 100) * 8)));
        CK(cudaFuncSetAttribute(marg_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(marg_eig_smem_doubles(256, 100) * 8)));
        ctx->attrs_set = true;
    }
#endif
    return 0;
}

enum { RUN_SOLVE = 1, RUN_REANCHOR = 2, RUN_MARG = 4, RUN_LIN_ONLY = 8 };

static int batch_execute(viwb_context *ctx, viwb_batch *b, int what) {
    bind_device(ctx);
    BatchDev &bd = b->bd;
    int rc = ensure_attrs(ctx); if (rc) return rc;
    stream_t st = ctx->stream;
    const int B = b->B;
    const size_t xs = b->total_state * sizeof(double);
    CK(dev_d2d(bd.x_cur, bd.x_init, xs, st)); CK(dev_d2d(bd.x_cand, bd.x_init, xs, st));
    if (!(what & RUN_REANCHOR) || (what & RUN_SOLVE)) CK(dev_d2d(bd.x_before, bd.x_init, xs, st));
    CK(dev_d2d(bd.work, b->work_init_dev, sizeof(WinWork) * B, st));
    const int nt_vis = NT(128), nt_lm = NT(128), nt_small = NT(128), nt_asm = NT(128), nt_syrk = NT(256), nt_solve = NT(SOLVE_NT), nt_marg = NT(256);
    const int g_vis = (bd.nvis_total + nt_vis - 1) / nt_vis, g_lm = (bd.nlm_total * LM_ROLES + nt_lm - 1) / nt_lm;
    const size_t sm_small = lin_small_smem_doubles(nt_small) * 8, sm_solve = solve_smem_doubles(nt_solve, bd.env_max) * 8, sm_marg = marg_prep_smem_doubles(nt_marg, bd.marg_nmax) * 8, sm_eig = marg_eig_smem_doubles(nt_marg, bd.marg_nmax) * 8;
    // cost_only: the round after the last allowed iteration only decides accept / reject of the pending candidate (every window
    // still running is at max_num_iterations there, trust_region_minimizer.cc checks the iteration limit before the gradient),
    // so the partial sums and the Schur product of that linearisation would never be read
    auto lin = [&](int mode, bool cost_only) {
        const bool solve = mode == MODE_SOLVE;
        const bool old_path = solve ? bd.n_unfused > 0 : bd.n_munfused > 0;      // windows outside the fused path (irregular factor tables)
        const int wpb2 = NT(128) / (NT(128) < 32 ? NT(128) : 32);
        const int nfi = solve ? bd.npitems_total : bd.nmpitems_total;               // fused pair items of this mode
        if (solve) {
            const int vmode = cost_only ? (int)MODE_COST : mode;
            if (bd.n_fused_compact > 0) { LAUNCH(lin_vis_lm, bd, bd.nlmb_total, 1, NT(LMB_FACTORS), lin_vis_lm_smem_doubles(false) * 8, vmode, st); ctx->launches++; }
            if (bd.n_fused_wide > 0) { LAUNCH(lin_vis_lm_wide, bd, bd.nlmb_total, 1, NT(LMB_FACTORS), lin_vis_lm_smem_doubles(true) * 8, vmode, st); ctx->launches++; }
        } else if (bd.n_mfused > 0) { LAUNCH(lin_vis_lm_wide, bd, bd.nlmb_total, 1, NT(LMB_FACTORS), lin_vis_lm_smem_doubles(true) * 8, mode, st); ctx->launches++; }
        auto fused_asm = [&]() {
            if (nfi <= 0) return;
            const int pw = solve ? bd.pwin_smem : bd.pwin_smem_marg;
            if (pw > 0) { LAUNCH(pair_win, bd, B, 1, NT(256), (size_t)pw, mode, st); ctx->launches++; return; }
            if (solve && bd.n_fused_compact > 0) { LAUNCH(asm_pairs, bd, (nfi + wpb2 - 1) / wpb2, 1, NT(128), 0, mode, st); ctx->launches++; }
            if (!solve || bd.n_fused_wide > 0) { LAUNCH(asm_pairs_wide, bd, (nfi + wpb2 - 1) / wpb2, 1, NT(128), 0, mode, st); ctx->launches++; }
            LAUNCH(pair_reduce, bd, B, 1, NT(PAIR_RED_NT), (size_t)bd.pitems_max * sizeof(int), mode, st); ctx->launches++;
        };
        if (!old_path) {
            LAUNCH(lin_small, bd, B, 1, nt_small, sm_small, mode, st); ctx->launches++;
            if (cost_only) return;
            fused_asm();
            if (g_syrk_dfma) LAUNCH(syrk, bd, B, 1, nt_syrk, syrk_smem_doubles() * 8, mode, st);
            else LAUNCH(syrk_mma, bd, B, 1, NT(SYRK_NT), syrk_mma_smem_doubles() * 8, mode, st);
            ctx->launches++;
            return;
        }
        LAUNCH(lin_vis, bd, g_vis, 1, nt_vis, lin_vis_smem_doubles(nt_vis, mode == MODE_SOLVE ? bd.rec_stride_solve : (int)VREC) * 8, mode, st);
        if ((mode == MODE_SOLVE ? bd.rec_stride_solve : (int)VREC) == VREC) LAUNCH(lm_reduce_wide, bd, g_lm, 1, nt_lm, 0, mode, st);
        else LAUNCH(lm_reduce, bd, g_lm, 1, nt_lm, 0, mode, st);
        LAUNCH(lin_small, bd, B, 1, nt_small, sm_small, mode, st);
        ctx->launches += (g_vis > 0) + (g_lm > 0) + 1;
        if (cost_only) return;
        const int ni = mode == MODE_SOLVE ? bd.nitems_solve : bd.nitems_marg, wpb = nt_asm / (nt_asm < 32 ? nt_asm : 32);
        const bool wide = (mode == MODE_SOLVE ? bd.rec_stride_solve : (int)VREC) == VREC;     // items carry the common columns
        fused_asm();
        if (wide) LAUNCH(asm_items_split, bd, (ni * ASM_SPLIT + wpb - 1) / wpb, 1, nt_asm, 0, mode, st);
        else LAUNCH(asm_items, bd, (ni + wpb - 1) / wpb, 1, nt_asm, 0, mode, st);
        if (g_syrk_dfma) LAUNCH(syrk, bd, B, 1, nt_syrk, syrk_smem_doubles() * 8, mode, st);
        else LAUNCH(syrk_mma, bd, B, 1, NT(SYRK_NT), syrk_mma_smem_doubles() * 8, mode, st);
        ctx->launches += (ni > 0) + 1;
    };
    if (what & (RUN_SOLVE | RUN_MARG | RUN_LIN_ONLY)) {
        const int ns = bd.nimu_total + bd.nwheel_total, nt_s = NT(128);
        LAUNCH(setup, bd, (ns + nt_s - 1) / nt_s, 1, nt_s, 0, 0, st);
        LAUNCH(prior_setup, bd, bd.nprior, 1, NT(256), 0, 0, st);
        ctx->launches += (ns > 0) + (bd.nprior > 0);
    }
    if (what & RUN_LIN_ONLY) { lin(MODE_SOLVE, false); LAUNCH(solve, bd, B, 1, nt_solve, sm_solve, 2, st); ctx->launches++; }
    if (what & RUN_SOLVE) {
        // SOLVER_TIME (estimator.cpp:1650-1653): Ceres tests the wall clock after every iteration; here the host
        // waits for each round only when a limit is set, and the round after the limit just decides the pending
        // candidate and stops (NO_CONVERGENCE).  Disabled (0) keeps the whole solve asynchronous.
        const auto t_start = std::chrono::steady_clock::now();
        bool last = false;
        for (int round = 0; round <= b->max_iter && !last; round++) {
            if (b->max_time > 0.0 && round > 0) {
                CK(dev_sync(st));
                last = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() >= b->max_time;
            }
            lin(MODE_SOLVE, last || (round == b->max_iter && round > 0));
            LAUNCH(solve, bd, B, 1, nt_solve, sm_solve, last ? 3 : 0, st);
            ctx->launches++;
        }
    }
    if (what & RUN_REANCHOR) { LAUNCH(reanchor, bd, B, 1, NT(32), 0, 0, st); ctx->launches++; }
    if ((what & RUN_MARG) && b->any_marg) {
        lin(MODE_MARG, false);
        LAUNCH(marg_prep, bd, B, 1, nt_marg, sm_marg, 0, st);
        if (g_marg_one_kernel) { LAUNCH(marg_eig, bd, B, 1, nt_marg, sm_eig, 0, st); ctx->launches += 2; }
        else {
            LAUNCH(marg_tri, bd, B, 1, NT(MARG_TRI_NT), sm_eig, 0, st);
            LAUNCH(marg_ql, bd, B, 1, NT(32), 2 * MAXPRI * 8, 0, st);
            LAUNCH(marg_apply, bd, B, 1, NT(128), marg_apply_smem_doubles(128, bd.marg_nmax) * 8, 0, st);
            ctx->launches += 4;
        }
    }
#ifndef VIWB_HOST_EMU
    CK((int)cudaGetLastError());
#endif
    return 0;
}

static int batch_fetch(viwb_context *ctx, viwb_batch *b, double *const *states, viwb_summary *summaries, viwb_prior *priors_out) {
    bind_device(ctx);
    BatchDev &bd = b->bd;
    const int B = b->B, nmax = b->prior_nmax;
    // staging layout inside the pinned slab
    char *hs = b->arena.host + b->out_off;
    size_t o = 0;
    double *x = (double *)(hs + o); o += align_up(b->total_state * 8);
    WinWork *work = (WinWork *)(hs + o); o += align_up(sizeof(WinWork) * B);
    int *hdr = (int *)(hs + o); o += align_up((size_t)B * (3 + 2 * NB) * 4);
    double *r = (double *)(hs + o); o += align_up((size_t)B * MAXPRI * 8);
    double *x0 = (double *)(hs + o); o += align_up((size_t)B * SFIX * 8);
    double *J = (double *)(hs + o);
    CK(dev_d2h(x, bd.x_cur, b->total_state * sizeof(double), ctx->stream));
    CK(dev_d2h(work, bd.work, sizeof(WinWork) * B, ctx->stream));
    const bool want_pr = priors_out && b->any_marg;
    if (want_pr) {
        CK(dev_d2h(hdr, bd.marg_hdr, (size_t)B * (3 + 2 * NB) * sizeof(int), ctx->stream));
        CK(dev_d2h(r, bd.marg_r, (size_t)B * MAXPRI * sizeof(double), ctx->stream));
        CK(dev_d2h(x0, bd.marg_x0, (size_t)B * SFIX * sizeof(double), ctx->stream));
        CK(dev_d2h(J, bd.marg_J, (size_t)B * nmax * nmax * 8, ctx->stream));     // window w's n x n block (row stride n) heads its nmax^2 slot
    }
    CK(dev_sync(ctx->stream));
    std::atomic<int> rc_a(0);
    parallel_for(B, [&](int w) {           // ~54 KB of prior per window: the unpack into the caller's buffers runs on the host pool
        if (states && states[w]) memcpy(states[w], x + b->meta[w].state_off, sizeof(double) * b->state_sizes[w]);
        if (summaries) {
            viwb_summary &sm = summaries[w]; const WinWork &ww = work[w];
            sm.termination_type = ww.term; sm.num_iterations = ww.num_iterations; sm.num_successful_steps = ww.successful; sm.num_linear_solves = ww.num_linear;
            sm.initial_cost = ww.initial_cost; sm.final_cost = ww.x_cost; sm.final_radius = ww.radius; sm.final_mu = ww.mu;
        }
        if (!priors_out) return;
        viwb_prior *out = &priors_out[w];
        const int mode = b->out_mode[w];
        if (mode == 0 && want_pr) {
            const int *h = hdr + (size_t)w * (3 + 2 * NB);
            out->valid = h[0]; out->n = h[1]; out->num_blocks = h[2];
            for (int i = 0; i < h[2]; i++) { out->block_id[i] = h[3 + i]; out->block_idx[i] = h[3 + NB + i]; }
            const int n = h[1];
            if (h[0]) {
                memcpy(out->J, J + (size_t)w * nmax * nmax, sizeof(double) * n * n);
                memcpy(out->r, r + (size_t)w * MAXPRI, sizeof(double) * n);
                memcpy(out->x0, x0 + (size_t)w * SFIX, sizeof(double) * SFIX);
            }
            if (work[w].marg_status < 0) rc_a = VIWB_ERR_NUMERIC;
        } else if (mode == 1) {
            const HostPrior &hp = b->in_prior[w];
            out->valid = 1; out->n = hp.n; out->num_blocks = hp.nb;
            memcpy(out->block_id, hp.block_id, sizeof hp.block_id); memcpy(out->block_idx, hp.block_idx, sizeof hp.block_idx);
            memcpy(out->x0, hp.x0.data(), sizeof(double) * SFIX); memcpy(out->J, hp.J.data(), sizeof(double) * hp.n * hp.n); memcpy(out->r, hp.r.data(), sizeof(double) * hp.n);
        } else { out->valid = 0; out->n = 0; out->num_blocks = 0; }
    });
    const int rc = rc_a;
    if (rc) return fail(ctx, rc, "marginalisation: kept dimension exceeds the shared-memory eigen solver");
    return 0;
}

static void batch_free(viwb_context *ctx, viwb_batch *b) {
    if (!b) return;
    if (ctx) dev_sync(ctx->stream);            // the upload of batch_build is not waited for there: nothing may still read the staging slab when it is released or reused
    if (b->arena_cached) { if (ctx) ctx->arena.busy = false; }
    else arena_release(b->arena);
    delete b;
}

// ====================================================================================== C ABI
extern "C" int viwb_create(int device, viwb_context **out) {
    if (!out) return VIWB_ERR_INVALID;
    viwb_context *ctx = new viwb_context();
    ctx->device = device; ctx->launches = 0; ctx->h2d_bytes = 0; ctx->attrs_set = false; ctx->stream = 0; ctx->own_stream = 0; ctx->lk1 = nullptr; ctx->det1 = nullptr;
#ifndef VIWB_HOST_EMU
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) { delete ctx; return VIWB_ERR_CUDA; }   // no CPU fallback: fail loudly
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return VIWB_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return VIWB_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
#endif
    *out = ctx;
    return VIWB_OK;
}
extern "C" void viwb_destroy(viwb_context *ctx) {
    if (!ctx) return;
    lk_batch_free(ctx->lk1);
    det_free(ctx->det1);
#ifndef VIWB_HOST_EMU
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
#endif
    arena_release(ctx->arena);
    delete ctx;
}
extern "C" const char *viwb_last_error(const viwb_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
// (NULL is CUDA's legacy default stream, a valid handle -- torch's current stream usually is it.)  Work already queued on the previous stream is
// waited for, so that buffers the context recycles (arena, staging slab) are never touched from two streams at once
extern "C" int viwb_set_stream(viwb_context *ctx, void *s) {
    if (!ctx) return VIWB_ERR_INVALID;
    bind_device(ctx);
    CK(dev_sync(ctx->stream));
    ctx->stream = (stream_t)s;
    return VIWB_OK;
}
extern "C" long long viwb_launch_count(const viwb_context *ctx) { return ctx ? ctx->launches : 0; }
extern "C" long long viwb_h2d_bytes(const viwb_context *ctx) { return ctx ? ctx->h2d_bytes : 0; }

extern "C" int viwb_set_profiling(viwb_context *ctx, int enable) {
    if (!ctx) return VIWB_ERR_INVALID;
#ifndef VIWB_HOST_EMU
    g_prof.collect(); g_prof.on = enable != 0;
    if (enable) { g_prof.names.clear(); g_prof.ms.clear(); g_prof.count.clear(); }
#else
    (void)enable;
#endif
    return VIWB_OK;
}
extern "C" int viwb_profile_count(viwb_context *ctx) {
    (void)ctx;
#ifndef VIWB_HOST_EMU
    g_prof.collect(); return (int)g_prof.names.size();
#else
    return 0;
#endif
}
extern "C" int viwb_profile_get(viwb_context *ctx, int idx, char *name, int name_cap, double *total_ms, long long *launches) {
    (void)ctx;
#ifndef VIWB_HOST_EMU
    if (idx < 0 || idx >= (int)g_prof.names.size()) return VIWB_ERR_INVALID;
    if (name && name_cap > 0) { strncpy(name, g_prof.names[idx].c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (total_ms) *total_ms = g_prof.ms[idx];
    if (launches) *launches = g_prof.count[idx];
    return VIWB_OK;
#else
    (void)idx; (void)name; (void)name_cap; (void)total_ms; (void)launches; return VIWB_ERR_INVALID;
#endif
}

extern "C" int viwb_batch_create(viwb_context *ctx, int batch, const viwb_problem *problems, const double *const *states,
                                 const viwb_options *options, const int32_t *margin_flags, viwb_batch **out) {
    if (!ctx || !out) return VIWB_ERR_INVALID;
    int rc = batch_build(ctx, batch, problems, states, options, margin_flags, out);
    if (rc) return rc;
    CK(dev_sync(ctx->stream));
    return VIWB_OK;
}
// the windows as uploaded: current and candidate states back to the initial ones, solver bookkeeping cleared (viwb_batch_run does the same before it starts;
// this entry lets a caller read / evaluate the pristine batch, e.g. viwb_batch_outliers on the initial guess)
extern "C" int viwb_batch_reset_states(viwb_context *ctx, viwb_batch *b) {
    if (!ctx || !b) return VIWB_ERR_INVALID;
    bind_device(ctx);
    const size_t xs = b->total_state * sizeof(double);
    CK(dev_d2d(b->bd.x_cur, b->bd.x_init, xs, ctx->stream)); CK(dev_d2d(b->bd.x_cand, b->bd.x_init, xs, ctx->stream)); CK(dev_d2d(b->bd.x_before, b->bd.x_init, xs, ctx->stream));
    CK(dev_d2d(b->bd.work, b->work_init_dev, sizeof(WinWork) * b->B, ctx->stream));
    return VIWB_OK;
}
extern "C" int viwb_batch_run(viwb_context *ctx, viwb_batch *b) {
    if (!ctx || !b) return VIWB_ERR_INVALID;
    return batch_execute(ctx, b, RUN_SOLVE | RUN_REANCHOR | RUN_MARG);
}
extern "C" int viwb_batch_download(viwb_context *ctx, viwb_batch *b, double *const *states, viwb_summary *summaries, viwb_prior *priors_out) {
    if (!ctx || !b) return VIWB_ERR_INVALID;
    return batch_fetch(ctx, b, states, summaries, priors_out);
}
// outliersRejection on the states the batch currently holds (after viwb_batch_run: the solved, re-anchored windows)
extern "C" int viwb_batch_outliers(viwb_context *ctx, viwb_batch *b, double focal_length, double threshold_px, uint8_t *const *outliers) {
    if (!ctx || !b || !outliers) return VIWB_ERR_INVALID;
    bind_device(ctx);
    BatchDev bd = b->bd; bd.out_focal = focal_length; bd.out_thresh = threshold_px;
    const int nt = NT(128);
    LAUNCH(outlier, bd, (bd.nlm_total + nt - 1) / nt, 1, nt, 0, 0, ctx->stream);
    ctx->launches += bd.nlm_total > 0;
    std::vector<int> h(bd.nlm_total > 0 ? bd.nlm_total : 1);
    CK(dev_d2h(h.data(), bd.lm_outlier, (size_t)bd.nlm_total * sizeof(int), ctx->stream));
    CK(dev_sync(ctx->stream));
    for (int w = 0; w < b->B; w++) if (outliers[w]) { const WinMeta &m = b->meta[w]; for (int k = 0; k < m.nlm; k++) outliers[w][k] = (uint8_t)h[m.lm_off + k]; }
    return VIWB_OK;
}
extern "C" double viwb_batch_algorithmic_bytes(const viwb_batch *b) { return b ? b->algorithmic_bytes : 0.0; }
extern "C" void viwb_batch_destroy(viwb_context *ctx, viwb_batch *b) { batch_free(ctx, b); }

static int run_once(viwb_context *ctx, int B, const viwb_problem *problems, double *const *states, const viwb_options *opt,
                    const int32_t *flags, int what, viwb_summary *summaries, viwb_prior *priors, const double *const *before) {
    viwb_batch *b = nullptr;
    int rc = batch_build(ctx, B, problems, (const double *const *)states, opt, flags, &b, true);
    if (rc) return rc;
    if (before) {   // gauge re-anchoring against an explicit pre-solve state
        std::vector<double> xb(b->total_state);
        for (int w = 0; w < B; w++) memcpy(xb.data() + b->meta[w].state_off, before[w], sizeof(double) * b->state_sizes[w]);
        rc = dev_h2d(b->bd.x_before, xb.data(), xb.size() * sizeof(double), ctx->stream);
        if (!rc) rc = dev_sync(ctx->stream);
        if (rc) { batch_free(ctx, b); return fail(ctx, VIWB_ERR_CUDA, "upload of state_before failed"); }
    }
    const double t0 = now_ms();
    rc = batch_execute(ctx, b, what);
    if (g_timing) { dev_sync(ctx->stream); fprintf(stderr, "[viwb] execute %.2f ms\n", now_ms() - t0); }
    const double t1 = now_ms();
    if (!rc) rc = batch_fetch(ctx, b, states, summaries, priors);
    if (g_timing) fprintf(stderr, "[viwb] fetch %.2f ms\n", now_ms() - t1);
    batch_free(ctx, b);
    return rc;
}

extern "C" int viwb_window_solve(viwb_context *ctx, const viwb_problem *problem, double *state, const viwb_options *options, viwb_summary *summary) {
    if (!ctx || !problem || !state) return VIWB_ERR_INVALID;
    double *sp[1] = {state};
    return run_once(ctx, 1, problem, sp, options, nullptr, RUN_SOLVE, summary, nullptr, nullptr);
}
extern "C" int viwb_gauge_reanchor(viwb_context *ctx, const viwb_problem *problem, const double *state_before, double *state) {
    if (!ctx || !problem || !state || !state_before) return VIWB_ERR_INVALID;
    double *sp[1] = {state}; const double *bp[1] = {state_before};
    return run_once(ctx, 1, problem, sp, nullptr, nullptr, RUN_REANCHOR, nullptr, nullptr, bp);
}
extern "C" int viwb_outlier_rejection(viwb_context *ctx, const viwb_problem *problem, const double *state, double focal_length, double threshold_px,
                                      uint8_t *outliers) {
    if (!ctx || !problem || !state || !outliers) return VIWB_ERR_INVALID;
    viwb_batch *b = nullptr;
    const double *sp[1] = {state};
    int rc = batch_build(ctx, 1, problem, sp, nullptr, nullptr, &b, true);
    if (rc) return rc;
    rc = batch_execute(ctx, b, 0);                     // x_cur <- the given state, nothing else
    uint8_t *op[1] = {outliers};
    if (!rc) rc = viwb_batch_outliers(ctx, b, focal_length, threshold_px, op);
    batch_free(ctx, b);
    return rc;
}
extern "C" int viwb_marginalize(viwb_context *ctx, const viwb_problem *problem, const double *state, int margin_flag, viwb_prior *prior_out) {
    if (!ctx || !problem || !state || !prior_out) return VIWB_ERR_INVALID;
    std::vector<double> tmp(state, state + SFIX + problem->num_landmarks);
    double *sp[1] = {tmp.data()}; int32_t fl[1] = {margin_flag};
    return run_once(ctx, 1, problem, sp, nullptr, fl, RUN_MARG, nullptr, prior_out, nullptr);
}
extern "C" int viwb_optimization(viwb_context *ctx, const viwb_problem *problem, double *state, const viwb_options *options, int margin_flag,
                                 viwb_summary *summary, viwb_prior *prior_out) {
    if (!ctx || !problem || !state) return VIWB_ERR_INVALID;
    double *sp[1] = {state}; int32_t fl[1] = {margin_flag};
    return run_once(ctx, 1, problem, sp, options, prior_out ? fl : nullptr, RUN_SOLVE | RUN_REANCHOR | (prior_out ? RUN_MARG : 0), summary, prior_out, nullptr);
}
extern "C" int viwb_optimization_batch(viwb_context *ctx, int batch, const viwb_problem *problems, double *const *states, const viwb_options *options,
                                       const int32_t *margin_flags, viwb_summary *summaries, viwb_prior *priors_out) {
    if (!ctx || !problems || !states) return VIWB_ERR_INVALID;
    return run_once(ctx, batch, problems, states, options, priors_out ? margin_flags : nullptr, RUN_SOLVE | RUN_REANCHOR | (priors_out ? RUN_MARG : 0),
                    summaries, priors_out, nullptr);
}

extern "C" int viwb_debug_normal_equations(viwb_context *ctx, const viwb_problem *problem, const double *state, double *H, double *g, double *lm, double *cost) {
    if (!ctx || !problem || !state) return VIWB_ERR_INVALID;
    viwb_batch *b = nullptr;
    const double *sp[1] = {state};
    int rc = batch_build(ctx, 1, problem, sp, nullptr, nullptr, &b);
    if (rc) return rc;
    rc = batch_execute(ctx, b, RUN_LIN_ONLY);
    if (rc) { batch_free(ctx, b); return rc; }
    const int N = problem->num_landmarks;
    std::vector<double> Hd((size_t)TFIX * (TFIX + 1) / 2), gd(TFIX), a(N + 1), gl(N + 1), W((size_t)(N + 1) * VSUB), lc(N + 1);
    std::vector<WinWork> ww(1);
    int e = dev_d2h(Hd.data(), b->bd.Hpk, Hd.size() * 8, ctx->stream);
    e |= dev_d2h(gd.data(), b->bd.gpk, gd.size() * 8, ctx->stream);
    e |= dev_d2h(a.data(), b->bd.lm_a, (size_t)N * 8, ctx->stream); e |= dev_d2h(gl.data(), b->bd.lm_g, (size_t)N * 8, ctx->stream);
    e |= dev_d2h(W.data(), b->bd.lm_W, (size_t)N * VSUB * 8, ctx->stream); e |= dev_d2h(lc.data(), b->bd.lm_cost, (size_t)N * 8, ctx->stream);
    e |= dev_d2h(ww.data(), b->bd.work, sizeof(WinWork), ctx->stream);
    e |= dev_sync(ctx->stream);
    if (e) { batch_free(ctx, b); return fail(ctx, VIWB_ERR_CUDA, "download failed"); }
    // formatting only: scatter the packed active triangle into the caller's fixed-layout arrays
    const WinMeta &m = b->meta[0];
    if (H) { memset(H, 0, sizeof(double) * TFIX * TFIX);
        std::vector<int> rp(m.nf + 1, 0);
        for (int i = 0; i < m.nf; i++) rp[i + 1] = rp[i] + i - m.efirst[i] + 1;      // skyline row pointers
        for (int ba = 0; ba < NB; ba++) for (int bb = 0; bb < NB; bb++) if (m.tcol[ba] >= 0 && m.tcol[bb] >= 0)
            for (int p = 0; p < blk_tsize(ba); p++) for (int q = 0; q < blk_tsize(bb); q++) {
                int ci = m.tcol[ba] + p, cj = m.tcol[bb] + q;
                if (ci < cj) std::swap(ci, cj);
                H[(blk_toff(ba) + p) * TFIX + blk_toff(bb) + q] = cj >= m.efirst[ci] ? Hd[(size_t)rp[ci] - m.efirst[ci] + cj] : 0.0;   // outside the envelope: structural zero
            } }
    if (g) { memset(g, 0, sizeof(double) * TFIX); for (int ba = 0; ba < NB; ba++) if (m.tcol[ba] >= 0) for (int p = 0; p < blk_tsize(ba); p++) g[blk_toff(ba) + p] = gd[m.tcol[ba] + p]; }
    double c = ww[0].small_cost;
    for (int k = 0; k < N; k++) {
        c += lc[k];
        if (lm) { lm[k * 82] = a[k]; lm[k * 82 + 1] = gl[k];
            for (int p = 0; p < VSUB; p++) {
                bool act = true;
                if (p < 66) act = m.tcol[p / 6] >= 0; else if (p < 72) act = m.tcol[BLK_EX0] >= 0; else if (p < 78) act = m.tcol[BLK_EX1] >= 0; else if (p == 78) act = m.tcol[BLK_TD] >= 0; else act = false;
                lm[k * 82 + 2 + p] = act ? W[(size_t)k * VSUB + p] : 0.0;
            } }
    }
    if (cost) *cost = c;
    batch_free(ctx, b);
    return VIWB_OK;
}

// -------------------------------------------------------------------------------------- factor level
struct EvalArgs { int type, want_j; const double *in; double *out; };
// in : globals(11) | consts(287) | params(7 blocks x 9)        out: residual(15) | tangent Jacobian (15 x 30)
VIWB_D void eval_factor_device(const EvalArgs &a) {
    const double *gl = a.in, *c = a.in + 11, *p = a.in + 11 + 287;
    double *r = a.out, *J = a.out + 15;
    const double *P[7]; for (int i = 0; i < 7; i++) P[i] = p + 9 * i;
    if (a.type <= 2) {
        VisOut o;
        if (a.type == 0) vis_eval(0, c, P[0], P[1], P[2], P[2], P[3][0], P[4][0], gl + 3, a.want_j != 0, o);
        else if (a.type == 1) vis_eval(1, c, P[0], P[1], P[2], P[3], P[4][0], P[5][0], gl + 3, a.want_j != 0, o);
        else vis_eval(2, c, P[0], P[0], P[0], P[1], P[2][0], P[3][0], gl + 3, a.want_j != 0, o);
        r[0] = o.r[0]; r[1] = o.r[1];
        if (a.want_j) { for (int k = 0; k < 12; k++) { J[k] = o.JA[k]; J[12 + k] = o.JB[k]; J[24 + k] = o.JE0[k]; J[36 + k] = o.JE1[k]; } J[48] = o.Jl[0]; J[49] = o.Jl[1]; J[50] = o.Jtd[0]; J[51] = o.Jtd[1]; }
    } else if (a.type == 3) {
        double S[225]; sqrt_info_upper(15, c + 62, S);
        imu_eval(c, S, gl, P[0], P[1], P[2], P[3], a.want_j != 0, r, J);
    } else if (a.type == 4) {
        double S[36]; sqrt_info_upper(6, c + 25, S);
        wheel_eval(c, S, P[0], P[1], P[2], P[3][0], P[4][0], P[5][0], P[6][0], a.want_j != 0, r, J);
    } else {
        plane_eval(gl + 7, P[0], P[1], P[2], P[3][0], a.want_j != 0, r, J);
    }
}
#ifndef VIWB_HOST_EMU
__global__ void eval_factor_kernel(EvalArgs a) { if (threadIdx.x == 0 && blockIdx.x == 0) eval_factor_device(a); }
#endif

extern "C" int viwb_factor_evaluate(viwb_context *ctx, int type, const viwb_globals *gl, const double *consts, const double *const *params,
                                    double *residuals, double **jacobians) {
    bind_device(ctx);
    if (!ctx || !gl || !params || !residuals || type < 0 || type > 5) return VIWB_ERR_INVALID;
    static const int nblk[6] = {5, 6, 4, 4, 7, 4};
    static const int sizes[6][7] = {{7, 7, 7, 1, 1, 0, 0}, {7, 7, 7, 7, 1, 1, 0}, {7, 7, 1, 1, 0, 0, 0}, {7, 9, 7, 9, 0, 0, 0}, {7, 7, 7, 1, 1, 1, 1}, {7, 7, 4, 1, 0, 0, 0}};
    static const int nres[6] = {2, 2, 2, 15, 6, 3};
    static const int ncons[6] = {12, 12, 12, 287, 78, 0};
    std::vector<double> in(11 + 287 + 63, 0.0), out(15 + 450, 0.0);
    for (int k = 0; k < 3; k++) in[k] = gl->G[k];
    for (int k = 0; k < 4; k++) in[3 + k] = gl->vis_sqrt_info[k];
    for (int k = 0; k < 3; k++) in[7 + k] = gl->plane_sqrt_info[k];
    in[10] = gl->huber_delta;
    if (ncons[type]) { if (!consts) return VIWB_ERR_INVALID; memcpy(in.data() + 11, consts, sizeof(double) * ncons[type]); }
    for (int i = 0; i < nblk[type]; i++) memcpy(in.data() + 11 + 287 + 9 * i, params[i], sizeof(double) * sizes[type][i]);
    double *din = nullptr, *dout = nullptr;
    CK(dev_malloc((void **)&din, in.size() * 8)); CK(dev_malloc((void **)&dout, out.size() * 8));
    CK(dev_h2d(din, in.data(), in.size() * 8, ctx->stream));
    EvalArgs a; a.type = type; a.want_j = jacobians ? 1 : 0; a.in = din; a.out = dout;
#ifdef VIWB_HOST_EMU
    eval_factor_device(a);
#else
    eval_factor_kernel<<<1, 32, 0, ctx->stream>>>(a);
#endif
    ctx->launches++;
    CK(dev_d2h(out.data(), dout, out.size() * 8, ctx->stream)); CK(dev_sync(ctx->stream));
    dev_free(din); dev_free(dout);
    for (int i = 0; i < nres[type]; i++) residuals[i] = out[i];
    if (jacobians) {   // formatting only: tangent columns -> the reference's row-major (rows x global size) blocks, last pose column zero
        const double *J = out.data() + 15;
        int ld = 0, tcol[7];
        if (type <= 2) {   // record layout A | B | E0 | E1 | l | td, 2 rows each
            const int off0[5] = {0, 12, 24, 48, 50}, off1[6] = {0, 12, 24, 36, 48, 50}, off2[4] = {24, 36, 48, 50};
            const int *off = type == 0 ? off0 : type == 1 ? off1 : off2;
            for (int i = 0; i < nblk[type]; i++) {
                if (!jacobians[i]) continue;
                const int gs = sizes[type][i], ts = gs == 7 ? 6 : 1;
                for (int r = 0; r < 2; r++) { for (int c2 = 0; c2 < ts; c2++) jacobians[i][r * gs + c2] = J[off[i] + r * ts + c2]; if (gs == 7) jacobians[i][r * gs + 6] = 0.0; }
            }
        } else {
            int c0 = 0;
            for (int i = 0; i < nblk[type]; i++) { tcol[i] = c0; const int gs = sizes[type][i]; c0 += gs == 7 ? 6 : gs == 4 ? 3 : gs; }
            ld = c0;
            for (int i = 0; i < nblk[type]; i++) {
                if (!jacobians[i]) continue;
                const int gs = sizes[type][i], ts = gs == 7 ? 6 : gs == 4 ? 3 : gs;
                for (int r = 0; r < nres[type]; r++) { for (int c2 = 0; c2 < ts; c2++) jacobians[i][r * gs + c2] = J[r * ld + tcol[i] + c2]; for (int c2 = ts; c2 < gs; c2++) jacobians[i][r * gs + c2] = 0.0; }
            }
        }
    }
    return VIWB_OK;
}

struct PriorEvalArgs { PriorDev p; const double *J, *r, *x0, *x; double *res; };
VIWB_D void prior_eval_device(const PriorEvalArgs &a, int tid, int nt, double *dx) {
    if (tid == 0) prior_dx(a.p, a.x, a.x0, dx);
    VIWB_SYNC();
    for (int i = tid; i < a.p.n; i += nt) { double s = a.r[i]; for (int k = 0; k < a.p.n; k++) s += a.J[i * a.p.n + k] * dx[k]; a.res[i] = s; }
}
#ifndef VIWB_HOST_EMU
__global__ void prior_eval_kernel(PriorEvalArgs a) { __shared__ double dx[MAXPRI]; prior_eval_device(a, threadIdx.x, blockDim.x, dx); }
#endif
extern "C" int viwb_prior_evaluate(viwb_context *ctx, const viwb_prior *prior, const double *state, double *residuals, double *jacobian) {
    bind_device(ctx);
    if (!ctx || !prior || !state || !residuals || prior->n <= 0 || prior->n > MAXPRI) return VIWB_ERR_INVALID;
    const int n = prior->n;
    double *d = nullptr;
    const size_t tot = (size_t)n * n + n + SFIX + SFIX + n;
    CK(dev_malloc((void **)&d, tot * 8));
    {   int e = dev_h2d(d, prior->J, (size_t)n * n * 8, ctx->stream);
        if (!e) e = dev_h2d(d + (size_t)n * n, prior->r, n * 8, ctx->stream);
        if (!e) e = dev_h2d(d + (size_t)n * n + n, prior->x0, SFIX * 8, ctx->stream);
        if (!e) e = dev_h2d(d + (size_t)n * n + n + SFIX, state, SFIX * 8, ctx->stream);
        if (e) { dev_free(d); return fail(ctx, VIWB_ERR_CUDA, std::string("prior upload: ") + dev_errstr(e)); } }
    PriorEvalArgs a; memset(&a, 0, sizeof a);
    a.p.n = n; a.p.nb = prior->num_blocks;
    for (int i = 0; i < prior->num_blocks; i++) { a.p.block_id[i] = prior->block_id[i]; a.p.block_idx[i] = prior->block_idx[i]; }
    a.J = d; a.r = d + (size_t)n * n; a.x0 = a.r + n; a.x = a.x0 + SFIX; a.res = d + (size_t)n * n + n + 2 * SFIX;
#ifdef VIWB_HOST_EMU
    { double dx[MAXPRI]; prior_eval_device(a, 0, 1, dx); }
#else
    prior_eval_kernel<<<1, 128, 0, ctx->stream>>>(a);
#endif
    ctx->launches++;
    {   int e = dev_d2h(residuals, a.res, n * 8, ctx->stream);
        if (!e) e = dev_sync(ctx->stream);
        dev_free(d);
        if (e) return fail(ctx, VIWB_ERR_CUDA, std::string("prior download: ") + dev_errstr(e)); }
    if (jacobian) {   // formatting only (marginalization_factor.cpp:381-394): J_lin columns at the block's state offset
        memset(jacobian, 0, sizeof(double) * n * SFIX);
        for (int i = 0; i < prior->num_blocks; i++) {
            const int bq = prior->block_id[i], ls = blk_msize(bq), idx = prior->block_idx[i], off = blk_off(bq);
            for (int r = 0; r < n; r++) for (int k = 0; k < ls; k++) jacobian[(size_t)r * SFIX + off + k] = prior->J[(size_t)r * n + idx + k];
        }
    }
    return VIWB_OK;
}

// -------------------------------------------------------------------------------------- pre-integration
template <typename Args, typename EmuFn>
static int preint_run(viwb_context *ctx, Args &a, int n, const int32_t *counts, const double *dt, const double *s0, const double *s1, int rec_doubles, double *records,
                      const double *const *extra, const int *extra_doubles, int nextra, const double **dev_extra, EmuFn emu, int smem_doubles, bool imu) {
    bind_device(ctx);
    std::vector<int> off(n + 1, 0);
    for (int i = 0; i < n; i++) { if (counts[i] < 0) return fail(ctx, VIWB_ERR_INVALID, "negative sample count"); off[i + 1] = off[i] + counts[i]; }
    const size_t steps = off[n], rows = steps + n;
    size_t bytes = align_up((n + 1) * sizeof(int)) + align_up(steps * 8) + 2 * align_up(rows * 24) + align_up((size_t)n * rec_doubles * 8);
    for (int k = 0; k < nextra; k++) bytes += align_up((size_t)n * extra_doubles[k] * 8);
    char *d = nullptr;
    CK(dev_malloc((void **)&d, bytes));
    size_t o = 0;
    int *d_off = (int *)(d + o); o += align_up((n + 1) * sizeof(int));
    double *d_dt = (double *)(d + o); o += align_up(steps * 8);
    double *d_s0 = (double *)(d + o); o += align_up(rows * 24);
    double *d_s1 = (double *)(d + o); o += align_up(rows * 24);
    double *d_rec = (double *)(d + o); o += align_up((size_t)n * rec_doubles * 8);
    int e = dev_h2d(d_off, off.data(), (n + 1) * sizeof(int), ctx->stream);
    if (!e) e = dev_h2d(d_dt, dt, steps * 8, ctx->stream);
    if (!e) e = dev_h2d(d_s0, s0, rows * 24, ctx->stream);
    if (!e) e = dev_h2d(d_s1, s1, rows * 24, ctx->stream);
    for (int k = 0; k < nextra && !e; k++) { double *p = (double *)(d + o); o += align_up((size_t)n * extra_doubles[k] * 8); e = dev_h2d(p, extra[k], (size_t)n * extra_doubles[k] * 8, ctx->stream); dev_extra[k] = p; }
    if (e) { dev_free(d); return fail(ctx, VIWB_ERR_CUDA, "pre-integration upload failed"); }
    a.n = n; a.off = d_off; a.dt = d_dt; a.rec = d_rec;
    emu(a, d_s0, d_s1);
    (void)smem_doubles; (void)imu;
    ctx->launches++;
    e = dev_d2h(records, d_rec, (size_t)n * rec_doubles * 8, ctx->stream);
    if (!e) e = dev_sync(ctx->stream);
    dev_free(d);
    if (e) return fail(ctx, VIWB_ERR_CUDA, "pre-integration download failed");
    return VIWB_OK;
}

extern "C" int viwb_imu_preintegrate(viwb_context *ctx, int n, const int32_t *counts, const double *dt, const double *acc, const double *gyr,
                                     const double *ba, const double *bg, const double *noise, double *records) {
    if (!ctx || n < 0 || !counts || !dt || !acc || !gyr || !ba || !bg || !noise || !records) return VIWB_ERR_INVALID;
    if (n == 0) return VIWB_OK;
    ImuPreArgs a; for (int k = 0; k < 4; k++) a.noise[k] = noise[k];
    const double *extra[2] = {ba, bg}; const int ed[2] = {3, 3}; const double *dv[2] = {nullptr, nullptr};
    stream_t st = ctx->stream;
    return preint_run(ctx, a, n, counts, dt, acc, gyr, 287, records, extra, ed, 2, dv, [&](ImuPreArgs &q, const double *d0, const double *d1) {
        q.acc = d0; q.gyr = d1; q.ba = dv[0]; q.bg = dv[1];
#ifdef VIWB_HOST_EMU
        std::vector<double> sm(PRE_IMU_SMEM); for (int i = 0; i < q.n; i++) imu_preint_warp(q, i, 0, 1, sm.data());
#else
        g_prof.begin("imu_preint", st); imu_preint_kernel<<<(q.n + PRE_WPB - 1) / PRE_WPB, 32 * PRE_WPB, PRE_WPB * PRE_IMU_SMEM * 8, st>>>(q); g_prof.end(st);
#endif
    }, PRE_IMU_SMEM, true);
}
extern "C" int viwb_wheel_preintegrate(viwb_context *ctx, int n, const int32_t *counts, const double *dt, const double *vel, const double *gyr,
                                       const double *s, const double *td, const double *noise, double *records) {
    if (!ctx || n < 0 || !counts || !dt || !vel || !gyr || !s || !td || !noise || !records) return VIWB_ERR_INVALID;
    if (n == 0) return VIWB_OK;
    WheelPreArgs a; a.noise[0] = noise[0]; a.noise[1] = noise[1];
    const double *extra[2] = {s, td}; const int ed[2] = {3, 1}; const double *dv[2] = {nullptr, nullptr};
    stream_t st = ctx->stream;
    return preint_run(ctx, a, n, counts, dt, vel, gyr, 78, records, extra, ed, 2, dv, [&](WheelPreArgs &q, const double *d0, const double *d1) {
        q.vel = d0; q.gyr = d1; q.s = dv[0]; q.td = dv[1];
#ifdef VIWB_HOST_EMU
        std::vector<double> sm(PRE_WHEEL_SMEM); for (int i = 0; i < q.n; i++) wheel_preint_warp(q, i, 0, 1, sm.data());
#else
        g_prof.begin("wheel_preint", st); wheel_preint_kernel<<<(q.n + PRE_WPB - 1) / PRE_WPB, 32 * PRE_WPB, PRE_WPB * PRE_WHEEL_SMEM * 8, st>>>(q); g_prof.end(st);
#endif
    }, PRE_WHEEL_SMEM, false);
}

// -------------------------------------------------------------------------------------- initialisation alignment (SURVEY 8 f-4 ii)
struct InitWork { char *d = nullptr; double *R, *T, *imu, *wheel, *A, *b, *x, *out; int *perm; };
static int init_work(viwb_context *ctx, InitWork &w, int F, const double *R, const double *T, const double *imu, const double *wheel) {
    const int n = 3 * F + 4;
    const size_t sR = align_up((size_t)F * 72), sT = align_up((size_t)F * 24), sI = align_up((size_t)(F - 1) * VIWB_IMU_DOUBLES * 8), sW = align_up((size_t)(F - 1) * VIWB_WHEEL_DOUBLES * 8);
    const size_t sA = align_up((size_t)2 * n * n * 8), sv = align_up((size_t)(n + 8) * 8);
    if (dev_malloc((void **)&w.d, sR + sT + sI + sW + sA + 3 * sv + align_up((size_t)n * 4))) return 1;
    size_t o = 0;
    w.R = (double *)(w.d + o); o += sR; w.T = (double *)(w.d + o); o += sT; w.imu = (double *)(w.d + o); o += sI; w.wheel = (double *)(w.d + o); o += sW;
    w.A = (double *)(w.d + o); o += sA; w.b = (double *)(w.d + o); o += sv; w.x = (double *)(w.d + o); o += sv; w.out = (double *)(w.d + o); o += sv;
    w.perm = (int *)(w.d + o);
    int e = dev_h2d(w.R, R, (size_t)F * 72, ctx->stream);
    if (!e && T) e = dev_h2d(w.T, T, (size_t)F * 24, ctx->stream);
    if (!e) e = dev_h2d(w.imu, imu, (size_t)(F - 1) * VIWB_IMU_DOUBLES * 8, ctx->stream);
    if (!e && wheel) e = dev_h2d(w.wheel, wheel, (size_t)(F - 1) * VIWB_WHEEL_DOUBLES * 8, ctx->stream);
    return e;
}

extern "C" int viwb_solve_gyroscope_bias(viwb_context *ctx, int num_frames, const double *R, const double *imu_data, double *delta_bg) {
    if (!ctx || !R || !imu_data || !delta_bg) return VIWB_ERR_INVALID;
    if (num_frames < 2 || num_frames > VIWB_MAX_INIT_FRAMES) return fail(ctx, VIWB_ERR_INVALID, "solve_gyroscope_bias: 2..VIWB_MAX_INIT_FRAMES frames");
    bind_device(ctx);
    InitWork w; int e = init_work(ctx, w, num_frames, R, nullptr, imu_data, nullptr);
    if (!e) {
        GyroBiasArgs a; a.F = num_frames; a.R = w.R; a.imu = w.imu; a.A = w.A; a.b = w.b; a.x = w.x; a.perm = w.perm; a.out = w.out;
#ifdef VIWB_HOST_EMU
        gyro_bias_block(a, 0, 1);
#else
        g_prof.begin("gyro_bias", ctx->stream); gyro_bias_kernel<<<1, 32, 0, ctx->stream>>>(a); g_prof.end(ctx->stream);
#endif
        ctx->launches++;
        e = dev_d2h(delta_bg, w.out, 24, ctx->stream);
    }
    if (!e) e = dev_sync(ctx->stream);
    dev_free(w.d);
    return e ? fail(ctx, VIWB_ERR_CUDA, "solve_gyroscope_bias failed") : VIWB_OK;
}

extern "C" int viwb_linear_alignment(viwb_context *ctx, int num_frames, const double *R, const double *T, const double *imu_data, const double *wheel_data,
                                     const double *tic, const double *rio, const double *tio, double g_norm, double *g, double *x, int32_t *x_size, int32_t *aligned) {
    if (!ctx || !R || !T || !imu_data || !tic || !g || !x || !x_size || !aligned || (wheel_data && (!rio || !tio))) return VIWB_ERR_INVALID;
    if (num_frames < 2 || num_frames > VIWB_MAX_INIT_FRAMES) return fail(ctx, VIWB_ERR_INVALID, "linear_alignment: 2..VIWB_MAX_INIT_FRAMES frames");
    bind_device(ctx);
    const int n = 3 * num_frames + 4;
    InitWork w; int e = init_work(ctx, w, num_frames, R, T, imu_data, wheel_data);
    std::vector<double> out((size_t)n + 8);
    if (!e) {
        AlignArgs a; a.F = num_frames; a.use_wheel = wheel_data ? 1 : 0; a.R = w.R; a.T = w.T; a.imu = w.imu; a.wheel = wheel_data ? w.wheel : nullptr;
        memcpy(a.tic, tic, 24); a.g_norm = g_norm;
        for (int k = 0; k < 9; k++) a.rio[k] = rio ? rio[k] : (k % 4 == 0 ? 1.0 : 0.0);
        for (int k = 0; k < 3; k++) a.tio[k] = tio ? tio[k] : 0.0;
        a.A = w.A; a.b = w.b; a.x = w.x; a.perm = w.perm; a.out = w.out;
#ifdef VIWB_HOST_EMU
        align_block(a, 0, 1);
#else
        g_prof.begin("align", ctx->stream); align_kernel<<<1, 128, 0, ctx->stream>>>(a); g_prof.end(ctx->stream);
#endif
        ctx->launches++;
        e = dev_d2h(out.data(), w.out, (size_t)(n + 5) * 8, ctx->stream);
    }
    if (!e) e = dev_sync(ctx->stream);
    dev_free(w.d);
    if (e) return fail(ctx, VIWB_ERR_CUDA, "linear_alignment failed");
    *aligned = out[0] != 0.0; memcpy(g, out.data() + 1, 24); *x_size = (int32_t)out[4];
    memcpy(x, out.data() + 5, (size_t)*x_size * 8);
    return VIWB_OK;
}

// -------------------------------------------------------------------------------------- triangulation / depth shift
extern "C" int viwb_triangulate(viwb_context *ctx, const double *state, int n, const int32_t *stereo, const int32_t *frame, const double *pt0, const double *pt1,
                                double init_depth, double *depth) {
    if (!ctx || !state || n < 0 || !stereo || !frame || !pt0 || !pt1 || !depth) return VIWB_ERR_INVALID;
    if (n == 0) return VIWB_OK;
    for (int k = 0; k < n; k++) if (frame[k] < 0 || frame[k] + (stereo[k] ? 0 : 1) > VIWB_WINDOW_SIZE) return fail(ctx, VIWB_ERR_INVALID, "triangulate: frame index out of the window");
    bind_device(ctx);
    const size_t bytes = align_up(SFIX * 8) + 2 * align_up((size_t)n * 4) + 2 * align_up((size_t)n * 16) + align_up((size_t)n * 8);
    char *d = nullptr; CK(dev_malloc((void **)&d, bytes));
    size_t o = 0;
    double *d_x = (double *)(d + o); o += align_up(SFIX * 8);
    int *d_st = (int *)(d + o); o += align_up((size_t)n * 4);
    int *d_fr = (int *)(d + o); o += align_up((size_t)n * 4);
    double *d_p0 = (double *)(d + o); o += align_up((size_t)n * 16);
    double *d_p1 = (double *)(d + o); o += align_up((size_t)n * 16);
    double *d_out = (double *)(d + o);
    int e = dev_h2d(d_x, state, SFIX * 8, ctx->stream);
    if (!e) e = dev_h2d(d_st, stereo, (size_t)n * 4, ctx->stream);
    if (!e) e = dev_h2d(d_fr, frame, (size_t)n * 4, ctx->stream);
    if (!e) e = dev_h2d(d_p0, pt0, (size_t)n * 16, ctx->stream);
    if (!e) e = dev_h2d(d_p1, pt1, (size_t)n * 16, ctx->stream);
    TriArgs a; a.n = n; a.state = d_x; a.stereo = d_st; a.frame = d_fr; a.pt0 = d_p0; a.pt1 = d_p1; a.init_depth = init_depth; a.depth = d_out;
    if (!e) {
#ifdef VIWB_HOST_EMU
        for (int k = 0; k < n; k++) triangulate_item(a, k);
#else
        g_prof.begin("triangulate", ctx->stream); triangulate_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(a); g_prof.end(ctx->stream);
#endif
        ctx->launches++;
        e = dev_d2h(depth, d_out, (size_t)n * 8, ctx->stream);
    }
    if (!e) e = dev_sync(ctx->stream);
    dev_free(d);
    return e ? fail(ctx, VIWB_ERR_CUDA, "triangulate failed") : VIWB_OK;
}
extern "C" int viwb_shift_depth(viwb_context *ctx, int n, const double *uv, const double *depth_in, const double *marg_R, const double *marg_P,
                                const double *new_R, const double *new_P, double init_depth, double *depth_out) {
    if (!ctx || n < 0 || !uv || !depth_in || !marg_R || !marg_P || !new_R || !new_P || !depth_out) return VIWB_ERR_INVALID;
    if (n == 0) return VIWB_OK;
    bind_device(ctx);
    const size_t bytes = align_up((size_t)n * 24) + 2 * align_up((size_t)n * 8);
    char *d = nullptr; CK(dev_malloc((void **)&d, bytes));
    double *d_uv = (double *)d, *d_in = (double *)(d + align_up((size_t)n * 24)), *d_out = (double *)(d + align_up((size_t)n * 24) + align_up((size_t)n * 8));
    int e = dev_h2d(d_uv, uv, (size_t)n * 24, ctx->stream);
    if (!e) e = dev_h2d(d_in, depth_in, (size_t)n * 8, ctx->stream);
    ShiftArgs a; a.n = n; a.uv = d_uv; a.depth_in = d_in; a.depth_out = d_out; a.init_depth = init_depth;
    memcpy(a.margR, marg_R, 72); memcpy(a.margP, marg_P, 24); memcpy(a.newR, new_R, 72); memcpy(a.newP, new_P, 24);
    if (!e) {
#ifdef VIWB_HOST_EMU
        for (int k = 0; k < n; k++) shift_depth_item(a, k);
#else
        g_prof.begin("shift_depth", ctx->stream); shift_depth_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(a); g_prof.end(ctx->stream);
#endif
        ctx->launches++;
        e = dev_d2h(depth_out, d_out, (size_t)n * 8, ctx->stream);
    }
    if (!e) e = dev_sync(ctx->stream);
    dev_free(d);
    return e ? fail(ctx, VIWB_ERR_CUDA, "shift_depth failed") : VIWB_OK;
}

extern "C" int viwb_undistort_velocity(viwb_context *ctx, const viwb_pinhole *cam, int n, const float *pts, const float *prev_un_pts, const uint8_t *has_prev,
                                       double dt, float *un_pts, float *velocity) {
    if (!ctx || !cam || n < 0 || !pts || !un_pts) return VIWB_ERR_INVALID;
    if (n == 0) return VIWB_OK;
    bind_device(ctx);
    const size_t pb = align_up((size_t)n * 8);
    char *d = nullptr; CK(dev_malloc((void **)&d, 4 * pb + align_up((size_t)n)));
    float *d_pts = (float *)d, *d_prev = (float *)(d + pb), *d_un = (float *)(d + 2 * pb), *d_vel = (float *)(d + 3 * pb);
    unsigned char *d_hp = (unsigned char *)(d + 4 * pb);
    int e = dev_h2d(d_pts, pts, (size_t)n * 8, ctx->stream);
    if (!e && prev_un_pts) e = dev_h2d(d_prev, prev_un_pts, (size_t)n * 8, ctx->stream);
    if (!e && has_prev) e = dev_h2d(d_hp, has_prev, (size_t)n, ctx->stream);
    UndistArgs a; a.n = n; a.pts = d_pts; a.prev_un = prev_un_pts ? d_prev : nullptr; a.has_prev = has_prev ? d_hp : nullptr;
    a.fx = cam->fx; a.fy = cam->fy; a.cx = cam->cx; a.cy = cam->cy; a.k1 = cam->k1; a.k2 = cam->k2; a.p1 = cam->p1; a.p2 = cam->p2; a.dt = dt;
    a.un = d_un; a.vel = velocity ? d_vel : nullptr;
    if (!e) {
#ifdef VIWB_HOST_EMU
        for (int k = 0; k < n; k++) undistort_item(a, k);
#else
        g_prof.begin("undistort", ctx->stream); undistort_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(a); g_prof.end(ctx->stream);
#endif
        ctx->launches++;
        e = dev_d2h(un_pts, d_un, (size_t)n * 8, ctx->stream);
        if (!e && velocity) e = dev_d2h(velocity, d_vel, (size_t)n * 8, ctx->stream);
    }
    if (!e) e = dev_sync(ctx->stream);
    dev_free(d);
    return e ? fail(ctx, VIWB_ERR_CUDA, "undistort failed") : VIWB_OK;
}

// -------------------------------------------------------------------------------------- feature tracker
extern "C" int viwb_lk_track(viwb_context *ctx, const uint8_t *prev_img, const uint8_t *next_img, int width, int height, int stride,
                             const float *prev_pts, float *next_pts, int n, int win_size, int max_level, int max_iter, float eps, int flags,
                             float min_eig_threshold, uint8_t *status, float *err) {
    if (!ctx || !prev_img || !next_img || !prev_pts || !next_pts || !status || n < 0 || width <= 0 || height <= 0) return VIWB_ERR_INVALID;
    if (win_size != 21) return fail(ctx, VIWB_ERR_UNSUPPORTED, "only the reference's 21x21 window is supported");
    if (max_level < 0 || max_level > 3) return fail(ctx, VIWB_ERR_UNSUPPORTED, "maxLevel must be 0..3");
    return lk_track_single(ctx, prev_img, next_img, width, height, stride, prev_pts, next_pts, n, max_level, max_iter, eps, flags, min_eig_threshold, status, err);
}

extern "C" int viwb_track_checked(viwb_context *ctx, const uint8_t *img_a, const uint8_t *img_b, int width, int height, int stride,
                                  const float *pts_a, float *pts_b, int n, int mode, int flow_back, uint8_t *status) {
    if (!ctx || !img_a || !img_b || !pts_a || !pts_b || !status || n < 0 || width <= 0 || height <= 0 || (mode != 0 && mode != 1)) return VIWB_ERR_INVALID;
    return lk_track_checked_single(ctx, img_a, img_b, width, height, stride, pts_a, pts_b, n, mode, flow_back, status);
}

extern "C" int viwb_lk_batch_create(viwb_context *ctx, int streams, int width, int height, int max_points, int stereo, int flow_back, viwb_lk_batch **out) {
    if (!ctx || !out || streams <= 0 || width <= 0 || height <= 0 || max_points <= 0) return VIWB_ERR_INVALID;
    return lk_batch_build(ctx, streams, width, height, max_points, stereo ? 1 : 0, flow_back ? 1 : 0, out);
}
extern "C" void viwb_lk_batch_destroy(viwb_lk_batch *b) { lk_batch_free(b); }
extern "C" int viwb_lk_batch_upload(viwb_lk_batch *b, const uint8_t *const *prev, const uint8_t *const *cur, const uint8_t *const *right, int stride,
                                    const float *prev_pts, const int32_t *n_prev, const float *stereo_pts, const int32_t *n_stereo) {
    if (!b || stride < b->w) return VIWB_ERR_INVALID;
    return lk_batch_upload(b, prev, cur, right, stride, prev_pts, n_prev, stereo_pts, n_stereo);
}
extern "C" int viwb_lk_batch_run(viwb_lk_batch *b) { return b ? lk_batch_execute(b, 3) : VIWB_ERR_INVALID; }
extern "C" int viwb_lk_batch_download(viwb_lk_batch *b, float *cur_pts, uint8_t *status, float *right_pts, uint8_t *status_right) {
    return b ? lk_batch_fetch(b, cur_pts, status, right_pts, status_right) : VIWB_ERR_INVALID;
}
extern "C" int viwb_set_mask(viwb_context *ctx, int width, int height, const float *pts, const int32_t *track_cnt, int n, int min_dist,
                             const uint8_t *base_mask, uint8_t *mask_out, int32_t *keep, int32_t *n_keep) {
    if (!ctx) return VIWB_ERR_INVALID;
    return det_set_mask(ctx, width, height, pts, track_cnt, n, min_dist, base_mask, mask_out, keep, n_keep);
}
extern "C" int viwb_good_features_to_track(viwb_context *ctx, const uint8_t *image, int width, int height, int stride, int max_corners, double quality_level,
                                           double min_distance, const uint8_t *mask, int mask_stride, float *corners, int capacity, int32_t *n_corners) {
    if (!ctx || stride < width || (mask && mask_stride < width)) return VIWB_ERR_INVALID;
    return det_good_features(ctx, image, width, height, stride, max_corners, quality_level, min_distance, mask, mask_stride, corners, capacity, n_corners);
}
extern "C" int viwb_detector_create(viwb_context *ctx, int streams, int width, int height, int max_pts, int min_dist, viwb_detector **out) {
    if (!ctx || !out) return VIWB_ERR_INVALID;
    return det_build(ctx, streams, width, height, max_pts, (double)min_dist, out);
}
extern "C" void viwb_detector_destroy(viwb_detector *d) { det_free(d); }
extern "C" int viwb_detector_detect(viwb_detector *d, const uint8_t *const *images, int stride, const viwb_lk_batch *resident, const uint8_t *const *base_masks,
                                    const float *pts, const int32_t *track_cnt, const int32_t *n_pts, int max_cnt, double quality_level,
                                    int32_t *keep, int32_t *n_keep, float *new_pts, int32_t *n_new, uint8_t *mask_out) {
    if (!d || !pts || !track_cnt || !n_pts || !keep || !n_keep || !new_pts || !n_new || (images && stride < d->w) || !(quality_level > 0.0)) return VIWB_ERR_INVALID;
    return det_detect(d, images, stride, resident, base_masks, pts, track_cnt, n_pts, max_cnt, quality_level, keep, n_keep, new_pts, n_new, mask_out);
}
extern "C" double viwb_detector_algorithmic_bytes(const viwb_detector *d) {
    if (!d) return 0.0;
    return (double)d->F * ((double)d->w * d->h + (double)d->maxn * (8 + 4 + 4 + 8));
}
extern "C" double viwb_lk_batch_algorithmic_bytes(const viwb_lk_batch *b) {
    if (!b) return 0.0;
    // per tick: read the new left (+right) level-0 images once, write their three coarser levels; points in/out
    double img = 0.0; for (int l = 0; l <= b->levels; l++) img += (double)b->lw[l] * b->lh[l];
    return (double)b->F * ((b->stereo ? 2.0 : 1.0) * img + (double)b->maxn * (b->stereo ? 2 : 1) * (8 + 8 + 1));
}
// page-lock caller-owned host buffers (camera frames) so that uploads run at full PCIe rate and asynchronously
extern "C" int viwb_tracker_create(viwb_context *ctx, int streams, int width, int height, const viwb_tracker_config *config, viwb_tracker **out) {
    if (!ctx || !out) return VIWB_ERR_INVALID;
    if (width < 3 || height < 3) return fail(ctx, VIWB_ERR_INVALID, "tracker: bad image size");
    return trk_build(ctx, streams, width, height, config, out);
}
extern "C" void viwb_tracker_destroy(viwb_tracker *t) { trk_free(t); }
extern "C" int viwb_tracker_track(viwb_tracker *t, double cur_time, const uint8_t *const *left, const uint8_t *const *right, int stride, const float *predict_pts,
                                  const uint8_t *has_prediction) {
    if (!t) return VIWB_ERR_INVALID;
    if (stride < t->w) return fail(t->ctx, VIWB_ERR_INVALID, "tracker: stride smaller than the image width");
    if (predict_pts && !has_prediction) return fail(t->ctx, VIWB_ERR_INVALID, "tracker: predict_pts without has_prediction flags");
    return trk_track(t, cur_time, left, right, stride, predict_pts, has_prediction);
}
extern "C" int viwb_tracker_download(viwb_tracker *t, int32_t *n_left, int32_t *ids, int32_t *track_cnt, float *feat, int32_t *n_right, int32_t *ids_right,
                                     float *feat_right) {
    return t ? trk_fetch(t, n_left, ids, track_cnt, feat, n_right, ids_right, feat_right) : VIWB_ERR_INVALID;
}
extern "C" double viwb_tracker_algorithmic_bytes(const viwb_tracker *t) {
    if (!t) return 0.0;
    const viwb_lk_batch *b = t->lk;
    double px = 0; for (int l = 1; l <= b->levels; l++) px += (double)b->lw[l] * b->lh[l];
    const double img = (double)t->w * t->h, cams = t->stereo ? 2.0 : 1.0;
    return (double)t->F * (cams * (img + px) + img + (double)t->maxn * (cams * (24 + 4) + 4));
}
extern "C" int viwb_host_register(viwb_context *ctx, void *ptr, size_t bytes) {
    bind_device(ctx);
    if (!ctx || !ptr) return VIWB_ERR_INVALID;
#ifndef VIWB_HOST_EMU
    CK((int)cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
#else
    (void)bytes;
#endif
    return VIWB_OK;
}
extern "C" int viwb_host_unregister(viwb_context *ctx, void *ptr) {
    bind_device(ctx);
    if (!ctx || !ptr) return VIWB_ERR_INVALID;
#ifndef VIWB_HOST_EMU
    CK((int)cudaHostUnregister(ptr));
#endif
    return VIWB_OK;
}

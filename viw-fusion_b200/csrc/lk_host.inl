// lk_host.inl -- host side of the LK tracker (included by viwb.cu after the device abstraction and the profiler).
// LkFrames = device state of F camera streams: three image slots per stream (two alternating left-camera slots so
// that the previous tick's image and its pyramid stay resident, one right-camera slot), pyramids, point / status
// buffers and the task tables that the kernels of kernels_lk.cuh index with blockIdx.y.

#ifdef VIWB_HOST_EMU
static void lk_launch_pyr(const PyrArgs *t, int items, int ntasks, stream_t) { for (int k = 0; k < ntasks; k++) for (int i = 0; i < items; i++) pyr_down_item(t[k], i); }
static void lk_launch_post(const PostArgs *t, int items, int ntasks, stream_t) { for (int k = 0; k < ntasks; k++) for (int i = 0; i < items; i++) lk_post_item(t[k], i); }
static void lk_launch_track(const LkArgs *t, int maxn, int ntasks, stream_t, const LkMaps *, bool) {
    std::vector<unsigned char> sm(lk_smem_bytes(1) + 64);
    for (int k = 0; k < ntasks; k++) for (int p = 0; p < maxn; p++) lk_track_warp(t[k], nullptr, p, 0, sm.data());
}
static int dev_h2d_2d(void *d, size_t dp, const void *h, size_t hp, size_t w, size_t rows, stream_t) { for (size_t r = 0; r < rows; r++) memcpy((char *)d + r * dp, (const char *)h + r * hp, w); return 0; }
#else
static void lk_launch_pyr(const PyrArgs *t, int items, int ntasks, stream_t s) {
    if (items <= 0 || ntasks <= 0) return;
    g_prof.begin("lk_pyr_down", s); pyr_down_tasks_kernel<<<dim3((items + 255) / 256, ntasks), 256, 0, s>>>(t); g_prof.end(s);
}
static void lk_launch_post(const PostArgs *t, int items, int ntasks, stream_t s) {
    if (items <= 0 || ntasks <= 0) return;
    g_prof.begin("lk_post", s); lk_post_tasks_kernel<<<dim3((items + 127) / 128, ntasks), 128, 0, s>>>(t); g_prof.end(s);
}
static void lk_launch_track(const LkArgs *t, int maxn, int ntasks, stream_t s, const LkMaps *maps, bool use_tma) {
    if (maxn <= 0 || ntasks <= 0) return;
    g_prof.begin("lk_track", s); lk_track_tasks_kernel<<<dim3((maxn + LK_PPB - 1) / LK_PPB, ntasks), 32 * LK_PPB, lk_smem_bytes(LK_PPB), s>>>(t, use_tma ? maps : nullptr); g_prof.end(s);
}
// One tensor map per pyramid level over the level's stacked images (u8, [LK_SLOTS * F * rows][width], row pitch a multiple of 16 bytes), box 48 bytes x
// 32 rows, no swizzle, zero fill outside; tiles are requested at 16-byte aligned byte columns (kernels_lk.cuh).  cuTensorMapEncodeTiled is a driver entry point: it is looked up at run time so that libviwb.so does not link libcuda
// (the library must still load -- and fail loudly in viwb_create -- on a box without a driver).
#include <cuda.h>
typedef CUresult (*viwb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int lk_encode_maps(LkMaps *out, uint8_t *const *level_base, const int *lw, const int *lh, const int *ls, int images) {
    static_assert(sizeof(CUtensorMap) == sizeof(out->opaque[0]), "CUtensorMap is 128 bytes");
    void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return 1;
    for (int l = 0; l < LK_MAXLVL; l++) {
        const cuuint64_t dims[2] = {(cuuint64_t)lw[l], (cuuint64_t)lh[l] * (cuuint64_t)images};
        const cuuint64_t strides[1] = {(cuuint64_t)ls[l]};
        const cuuint32_t box[2] = {LK_JS, LK_JROWS}, estr[2] = {1, 1};
        CUtensorMap tm;
        const CUresult r = ((viwb_encode_tiled_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, level_base[l], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 2;
        memcpy(&out->opaque[l][0], &tm, sizeof tm);
    }
    return 0;
}
static int dev_h2d_2d(void *d, size_t dp, const void *h, size_t hp, size_t w, size_t rows, stream_t s) { return (w && rows) ? (int)cudaMemcpy2DAsync(d, dp, h, hp, w, rows, cudaMemcpyHostToDevice, s) : 0; }
#endif

// number of usable levels: buildOpticalFlowPyramid stops once a level is not larger than the window
static int lk_levels(int w, int h, int max_level) {
    int lv = 0, cw = w, ch = h;
    for (int l = 1; l <= max_level; l++) { cw = (cw + 1) / 2; ch = (ch + 1) / 2; if (cw <= LK_WIN || ch <= LK_WIN) break; lv = l; }
    return lv;
}
static void lk_criteria(int max_iter, float eps, int &mi, float &e2) {
    mi = max_iter < 0 ? 0 : (max_iter > 100 ? 100 : max_iter);
    float e = eps < 0.f ? 0.f : (eps > 10.f ? 10.f : eps);
    e2 = e * e;
}

enum { LK_SLOTS = 3, LK_PTS = 6, LK_STS = 4 };
struct viwb_lk_batch {
    viwb_context *ctx;
    int F, w, h, maxn, levels, stereo, flow_back;
    int lw[LK_MAXLVL], lh[LK_MAXLVL], ls[LK_MAXLVL]; size_t lsz[LK_MAXLVL];
    uint8_t *img[LK_MAXLVL];          // level l: [LK_SLOTS][F] images of lsz[l] bytes
    int cur;                          // left slot holding the current image (0/1); 1 - cur holds the previous one; slot 2 = right
    bool dirty[LK_SLOTS];             // slot was uploaded and its pyramid is not built yet
    float *pts;                       // [LK_PTS][F][maxn][2]: prev, cur(out), back, stereo in, right(out), back2
    uint8_t *st;                      // [LK_STS][F][maxn]: temporal, temporal back, stereo, stereo back
    float *err;                       // [LK_STS][F][maxn]
    int *cnt;                         // [2][F] point counts: temporal, stereo
    LkArgs *tasks;                    // [2 (cur)][2 (wave)][2F]
    PyrArgs *pyr;                     // [LK_SLOTS][levels][F]
    PostArgs *post;                   // [2F]
    LkArgs *single;                   // scratch task for the single-call entry points
    size_t bytes_images, bytes_points;
    LkMaps *maps; bool use_tma;       // device copy of the level stacks' tensor maps (interior windows of lk_track are staged by TMA)
    uint8_t *image(int l, int slot, int f) const { return img[l] + ((size_t)slot * F + f) * lsz[l]; }
    float *P(int k, int f) const { return pts + ((size_t)k * F + f) * maxn * 2; }
    uint8_t *S(int k, int f) const { return st + ((size_t)k * F + f) * maxn; }
    float *E(int k, int f) const { return err + ((size_t)k * F + f) * maxn; }
};

static void lk_fill_image(const viwb_lk_batch *b, LkImage &im, int slot, int f) {
    for (int l = 0; l < LK_MAXLVL; l++) { im.img[l] = b->image(l, slot, f); im.w[l] = b->lw[l]; im.h[l] = b->lh[l]; im.stride[l] = b->ls[l]; }
}
// rows of (slot, stream) inside the level stacks: the TMA coordinates of a task's template (I) and search (J) images
static void lk_fill_rows(const viwb_lk_batch *b, LkArgs &a, int slot_i, int slot_j, int f) {
    for (int l = 0; l < LK_MAXLVL; l++) { a.trowI[l] = (slot_i * b->F + f) * b->lh[l]; a.trowJ[l] = (slot_j * b->F + f) * b->lh[l]; }
    a.tma = b->use_tma ? 1 : 0;
}

static void lk_batch_free(viwb_lk_batch *b) {
    if (!b) return;
    for (int l = 0; l < LK_MAXLVL; l++) if (b->img[l]) dev_free(b->img[l]);
    if (b->pts) dev_free(b->pts); if (b->st) dev_free(b->st); if (b->err) dev_free(b->err); if (b->cnt) dev_free(b->cnt);
    if (b->maps) dev_free(b->maps);
    if (b->tasks) dev_free(b->tasks); if (b->pyr) dev_free(b->pyr); if (b->post) dev_free(b->post); if (b->single) dev_free(b->single);
    delete b;
}

static int lk_batch_build(viwb_context *ctx, int F, int w, int h, int maxn, int stereo, int flow_back, viwb_lk_batch **out) {
    bind_device(ctx);
    viwb_lk_batch *b = new viwb_lk_batch();
    memset(b, 0, sizeof *b);
    b->ctx = ctx; b->F = F; b->w = w; b->h = h; b->maxn = maxn; b->stereo = stereo; b->flow_back = flow_back; b->levels = lk_levels(w, h, 3); b->cur = 0;
    int cw = w, ch = h;
    for (int l = 0; l < LK_MAXLVL; l++) { b->lw[l] = cw; b->lh[l] = ch; b->ls[l] = (cw + 15) & ~15; b->lsz[l] = (size_t)b->ls[l] * ch; cw = (cw + 1) / 2; ch = (ch + 1) / 2; }
#define LKA(p, n) do { if (dev_malloc((void **)&(p), (n))) { lk_batch_free(b); return fail(ctx, VIWB_ERR_CUDA, "LK device allocation failed"); } } while (0)
    for (int l = 0; l < LK_MAXLVL; l++) { LKA(b->img[l], b->lsz[l] * LK_SLOTS * F); b->bytes_images += b->lsz[l] * LK_SLOTS * F; }
    LKA(b->pts, (size_t)LK_PTS * F * maxn * 8); LKA(b->st, (size_t)LK_STS * F * maxn); LKA(b->err, (size_t)LK_STS * F * maxn * 4); LKA(b->cnt, (size_t)2 * F * 4);
    LKA(b->tasks, sizeof(LkArgs) * 8 * F); LKA(b->pyr, sizeof(PyrArgs) * LK_SLOTS * 3 * F); LKA(b->post, sizeof(PostArgs) * 2 * F); LKA(b->single, sizeof(LkArgs));
#undef LKA
    b->bytes_points = (size_t)LK_PTS * F * maxn * 8;
    b->use_tma = false;
#ifndef VIWB_HOST_EMU
    if (!getenv("VIWB_LK_NO_TMA")) {
        if ((size_t)LK_SLOTS * F * h > 0x7fffffffull) { lk_batch_free(b); return fail(ctx, VIWB_ERR_INVALID, "LK batch too large for one tensor map"); }
        LkMaps hm;
        const int e = lk_encode_maps(&hm, b->img, b->lw, b->lh, b->ls, LK_SLOTS * F);
        if (e) { lk_batch_free(b); return fail(ctx, VIWB_ERR_CUDA, e == 1 ? "cuTensorMapEncodeTiled not available from this driver" : "cuTensorMapEncodeTiled failed"); }
        if (dev_malloc((void **)&b->maps, sizeof(LkMaps)) || cudaMemcpy(b->maps, &hm, sizeof hm, cudaMemcpyHostToDevice) != cudaSuccess) { lk_batch_free(b); return fail(ctx, VIWB_ERR_CUDA, "tensor map upload failed"); }
        b->use_tma = true;
    }
#endif
    // ---- task tables (they only hold addresses inside this object, so they are built once)
    std::vector<LkArgs> T((size_t)8 * F); std::vector<PyrArgs> Y((size_t)LK_SLOTS * 3 * F); std::vector<PostArgs> Q((size_t)2 * F);
    int mi; float e2; lk_criteria(30, 0.01f, mi, e2);
    for (int cur = 0; cur < 2; cur++) for (int f = 0; f < F; f++) {
        const int prev = 1 - cur;
        LkArgs *w1 = T.data() + ((size_t)cur * 2 + 0) * 2 * F, *w2 = T.data() + ((size_t)cur * 2 + 1) * 2 * F;
        LkArgs a; memset(&a, 0, sizeof a); a.max_iter = mi; a.eps2 = e2; a.min_eig = 1e-4f;
        // temporal forward: prev -> cur, maxLevel 3 (feature_tracker.cpp:139)
        lk_fill_image(b, a.I, prev, f); lk_fill_image(b, a.J, cur, f); lk_fill_rows(b, a, prev, cur, f);
        a.prev_pts = b->P(0, f); a.next_pts = b->P(1, f); a.status = b->S(0, f); a.err = b->E(0, f); a.n_dev = b->cnt + f; a.max_level = b->levels; a.flags = 0;
        w1[f] = a;
        // temporal reverse: cur -> prev, maxLevel 1, OPTFLOW_USE_INITIAL_FLOW seeded with prev_pts (:144-146)
        lk_fill_image(b, a.I, cur, f); lk_fill_image(b, a.J, prev, f); lk_fill_rows(b, a, cur, prev, f);
        a.prev_pts = b->P(1, f); a.next_pts = b->P(2, f); a.status = b->S(1, f); a.err = b->E(1, f); a.max_level = b->levels < 1 ? b->levels : 1; a.flags = 4;
        w2[f] = a;
        // stereo forward: cur -> right, maxLevel 3 (:240)
        lk_fill_image(b, a.I, cur, f); lk_fill_image(b, a.J, 2, f); lk_fill_rows(b, a, cur, 2, f);
        a.prev_pts = b->P(3, f); a.next_pts = b->P(4, f); a.status = b->S(2, f); a.err = b->E(2, f); a.n_dev = b->cnt + F + f; a.max_level = b->levels; a.flags = 0;
        w1[F + f] = a;
        // stereo reverse: right -> cur, maxLevel 3, no initial flow (:244)
        lk_fill_image(b, a.I, 2, f); lk_fill_image(b, a.J, cur, f); lk_fill_rows(b, a, 2, cur, f);
        a.prev_pts = b->P(4, f); a.next_pts = b->P(5, f); a.status = b->S(3, f); a.err = b->E(3, f);
        w2[F + f] = a;
    }
    for (int s = 0; s < LK_SLOTS; s++) for (int l = 1; l <= 3; l++) for (int f = 0; f < F; f++) {
        PyrArgs p; p.src = b->image(l - 1, s, f); p.dst = b->image(l, s, f); p.sw = b->lw[l - 1]; p.sh = b->lh[l - 1]; p.sstride = b->ls[l - 1]; p.dw = b->lw[l]; p.dh = b->lh[l]; p.dstride = b->ls[l];
        Y[((size_t)s * 3 + (l - 1)) * F + f] = p;
    }
    for (int f = 0; f < F; f++) {
        PostArgs p; memset(&p, 0, sizeof p); p.w = w; p.h = h; p.flow_back = flow_back;
        p.pts_a = b->P(0, f); p.pts_b = b->P(1, f); p.pts_back = b->P(2, f); p.status = b->S(0, f); p.status_back = b->S(1, f); p.n_dev = b->cnt + f; p.mode = 0; Q[f] = p;
        p.pts_a = b->P(3, f); p.pts_b = b->P(4, f); p.pts_back = b->P(5, f); p.status = b->S(2, f); p.status_back = b->S(3, f); p.n_dev = b->cnt + F + f; p.mode = 1; Q[F + f] = p;
    }
    int e = dev_h2d(b->tasks, T.data(), sizeof(LkArgs) * T.size(), ctx->stream);
    if (!e) e = dev_h2d(b->pyr, Y.data(), sizeof(PyrArgs) * Y.size(), ctx->stream);
    if (!e) e = dev_h2d(b->post, Q.data(), sizeof(PostArgs) * Q.size(), ctx->stream);
    if (!e) e = dev_sync(ctx->stream);
    if (e) { lk_batch_free(b); return fail(ctx, VIWB_ERR_CUDA, "LK task upload failed"); }
    *out = b;
    return VIWB_OK;
}

// F images (host pointer per stream) into one slot; one strided copy when the host images are equally spaced
static int lk_upload_slot(viwb_lk_batch *b, int slot, const uint8_t *const *imgs, int stride) {
    viwb_context *ctx = b->ctx;
    bind_device(ctx);
    const int F = b->F;
    bool spaced = stride == b->ls[0];
    const ptrdiff_t gap = F > 1 ? imgs[1] - imgs[0] : (ptrdiff_t)b->lsz[0];
    if (gap < (ptrdiff_t)b->lsz[0]) spaced = false;
    for (int f = 0; f < F && spaced; f++) if (imgs[f] - imgs[0] != gap * f) spaced = false;
    if (spaced) CK(dev_h2d_2d(b->image(0, slot, 0), b->lsz[0], imgs[0], (size_t)gap, b->lsz[0], F, ctx->stream));
    else for (int f = 0; f < F; f++) CK(dev_h2d_2d(b->image(0, slot, f), b->ls[0], imgs[f], stride, b->w, b->h, ctx->stream));
    b->dirty[slot] = true;
    return VIWB_OK;
}

static int lk_batch_upload(viwb_lk_batch *b, const uint8_t *const *prev, const uint8_t *const *cur, const uint8_t *const *right, int stride,
                           const float *prev_pts, const int32_t *n_prev, const float *stereo_pts, const int32_t *n_stereo) {
    viwb_context *ctx = b->ctx;
    const int F = b->F;
    if (cur) {                              // a new tick: the old current image becomes the previous one (feature_tracker.cpp:296)
        b->cur = 1 - b->cur;
        int rc = lk_upload_slot(b, b->cur, cur, stride); if (rc) return rc;
    }
    if (prev) { int rc = lk_upload_slot(b, 1 - b->cur, prev, stride); if (rc) return rc; }
    if (right && b->stereo) { int rc = lk_upload_slot(b, 2, right, stride); if (rc) return rc; }
    if (prev_pts) CK(dev_h2d(b->P(0, 0), prev_pts, (size_t)F * b->maxn * 8, ctx->stream));
    if (n_prev) CK(dev_h2d(b->cnt, n_prev, (size_t)F * 4, ctx->stream));
    if (stereo_pts && b->stereo) CK(dev_h2d(b->P(3, 0), stereo_pts, (size_t)F * b->maxn * 8, ctx->stream));
    if (n_stereo && b->stereo) CK(dev_h2d(b->cnt + F, n_stereo, (size_t)F * 4, ctx->stream));
    return VIWB_OK;
}

// what: bit 0 temporal, bit 1 stereo
// rebuild_cur = false: the current image's pyramid is only built if the slot is dirty (the session tracker runs the temporal and the
// stereo half of one tick as two calls)
static int lk_batch_execute(viwb_lk_batch *b, int what, bool rebuild_cur = true) {
    viwb_context *ctx = b->ctx;
    bind_device(ctx);
    const int F = b->F; stream_t st = ctx->stream;
    if (!b->stereo) what &= 1;
    // every tick brings new cur (and right) images, so their pyramids are part of the tick; the previous image keeps the
    // pyramid it got when it was the current one unless it was (re)uploaded
    for (int s = 0; s < LK_SLOTS; s++) {
        const bool need = (s == b->cur && (rebuild_cur || b->dirty[s])) || (s == 2 && (what & 2)) || (s == 1 - b->cur && (what & 1) && b->dirty[s]);
        if (!need) continue;
        for (int l = 1; l <= b->levels; l++) { lk_launch_pyr(b->pyr + ((size_t)s * 3 + (l - 1)) * F, pyr_items_wh(b->lw[l], b->lh[l]), F, st); ctx->launches++; }
        b->dirty[s] = false;
    }
    const LkArgs *w1 = b->tasks + ((size_t)b->cur * 2 + 0) * 2 * F, *w2 = b->tasks + ((size_t)b->cur * 2 + 1) * 2 * F;
    const int first = (what & 1) ? 0 : F, count = ((what & 1) ? F : 0) + ((what & 2) ? F : 0);
    if (count == 0) return VIWB_OK;
    // forward flows start from the source points (no OPTFLOW_USE_INITIAL_FLOW), the temporal reverse flow from prev_pts
    if (b->flow_back && (what & 1)) CK(dev_d2d(b->P(2, 0), b->P(0, 0), (size_t)F * b->maxn * 8, st));
    lk_launch_track(w1 + first, b->maxn, count, st, b->maps, b->use_tma); ctx->launches++;
    if (b->flow_back) { lk_launch_track(w2 + first, b->maxn, count, st, b->maps, b->use_tma); ctx->launches++; }
    lk_launch_post(b->post + first, b->maxn, count, st); ctx->launches++;
#ifndef VIWB_HOST_EMU
    CK((int)cudaGetLastError());
#endif
    return VIWB_OK;
}

static int lk_batch_fetch(viwb_lk_batch *b, float *cur_pts, uint8_t *status, float *right_pts, uint8_t *status_right) {
    viwb_context *ctx = b->ctx;
    bind_device(ctx);
    const size_t np = (size_t)b->F * b->maxn;
    if (cur_pts) CK(dev_d2h(cur_pts, b->P(1, 0), np * 8, ctx->stream));
    if (status) CK(dev_d2h(status, b->S(0, 0), np, ctx->stream));
    if (right_pts && b->stereo) CK(dev_d2h(right_pts, b->P(4, 0), np * 8, ctx->stream));
    if (status_right && b->stereo) CK(dev_d2h(status_right, b->S(2, 0), np, ctx->stream));
    CK(dev_sync(ctx->stream));
    return VIWB_OK;
}

// ---- single-call entry points: a one-stream LkFrames cached in the context
static int lk_single_frames(viwb_context *ctx, int w, int h, int n, viwb_lk_batch **out);

static int lk_track_single(viwb_context *ctx, const uint8_t *prev, const uint8_t *next, int w, int h, int stride, const float *prev_pts, float *next_pts, int n,
                           int max_level, int max_iter, float eps, int flags, float min_eig, uint8_t *status, float *err) {
    bind_device(ctx);
    if (n == 0) return VIWB_OK;
    viwb_lk_batch *b; int rc = lk_single_frames(ctx, w, h, n, &b); if (rc) return rc;
    const uint8_t *pa[1] = {prev}, *pb[1] = {next};
    rc = lk_upload_slot(b, 0, pa, stride); if (rc) return rc;
    rc = lk_upload_slot(b, 1, pb, stride); if (rc) return rc;
    stream_t st = ctx->stream;
    for (int s = 0; s < 2; s++) { for (int l = 1; l <= b->levels; l++) { lk_launch_pyr(b->pyr + ((size_t)s * 3 + (l - 1)) * b->F, pyr_items_wh(b->lw[l], b->lh[l]), 1, st); ctx->launches++; } b->dirty[s] = false; }
    CK(dev_h2d(b->P(0, 0), prev_pts, (size_t)n * 8, st));
    CK(dev_h2d(b->P(1, 0), (flags & 4) ? next_pts : prev_pts, (size_t)n * 8, st));
    LkArgs a; memset(&a, 0, sizeof a);
    lk_fill_image(b, a.I, 0, 0); lk_fill_image(b, a.J, 1, 0); lk_fill_rows(b, a, 0, 1, 0);
    a.prev_pts = b->P(0, 0); a.next_pts = b->P(1, 0); a.status = b->S(0, 0); a.err = b->E(0, 0); a.n = n; a.n_dev = nullptr;
    a.max_level = max_level < b->levels ? max_level : b->levels; a.flags = flags; a.min_eig = min_eig;
    lk_criteria(max_iter, eps, a.max_iter, a.eps2);
    CK(dev_h2d(b->single, &a, sizeof a, st));
#ifndef VIWB_HOST_EMU
    CK(dev_sync(st));                      // `a` lives on this stack frame
#endif
    lk_launch_track(b->single, n, 1, st, b->maps, b->use_tma); ctx->launches++;
    CK(dev_d2h(next_pts, b->P(1, 0), (size_t)n * 8, st));
    CK(dev_d2h(status, b->S(0, 0), (size_t)n, st));
    if (err) CK(dev_d2h(err, b->E(0, 0), (size_t)n * 4, st));
    CK(dev_sync(st));
    return VIWB_OK;
}

static int lk_single_frames(viwb_context *ctx, int w, int h, int n, viwb_lk_batch **out) {
    viwb_lk_batch *b = ctx->lk1;
    if (b && (b->w != w || b->h != h || b->maxn < n)) { lk_batch_free(b); ctx->lk1 = b = nullptr; }
    if (!b) {
        int cap = 256; while (cap < n) cap *= 2;
        int rc = lk_batch_build(ctx, 1, w, h, cap, 1, 1, &b); if (rc) return rc;
        ctx->lk1 = b;
    }
    *out = b;
    return VIWB_OK;
}

// forward + (optional) reverse LK sharing the two pyramids, then the reference's status rules, all on the device
static int lk_track_checked_single(viwb_context *ctx, const uint8_t *img_a, const uint8_t *img_b, int w, int h, int stride, const float *pts_a, float *pts_b,
                                   int n, int mode, int flow_back, uint8_t *status) {
    bind_device(ctx);
    if (n == 0) return VIWB_OK;
    viwb_lk_batch *b; int rc = lk_single_frames(ctx, w, h, n, &b); if (rc) return rc;
    if (b->flow_back != (flow_back ? 1 : 0)) {          // the status rule is baked into the task tables: rebuild them for the other setting
        std::vector<PostArgs> Q(2);
        CK(dev_d2h(Q.data(), b->post, sizeof(PostArgs) * 2, ctx->stream)); CK(dev_sync(ctx->stream));
        Q[0].flow_back = Q[1].flow_back = flow_back ? 1 : 0;
        CK(dev_h2d(b->post, Q.data(), sizeof(PostArgs) * 2, ctx->stream)); CK(dev_sync(ctx->stream));
        b->flow_back = flow_back ? 1 : 0;
    }
    const uint8_t *pa[1] = {img_a}, *pb[1] = {img_b};
    const int32_t cnt = n;
    stream_t st = ctx->stream;
    if (mode == 0) {
        rc = lk_upload_slot(b, 1 - b->cur, pa, stride); if (rc) return rc;
        rc = lk_upload_slot(b, b->cur, pb, stride); if (rc) return rc;
        CK(dev_h2d(b->P(0, 0), pts_a, (size_t)n * 8, st)); CK(dev_h2d(b->cnt, &cnt, 4, st));
    } else {
        rc = lk_upload_slot(b, b->cur, pa, stride); if (rc) return rc;
        rc = lk_upload_slot(b, 2, pb, stride); if (rc) return rc;
        CK(dev_h2d(b->P(3, 0), pts_a, (size_t)n * 8, st)); CK(dev_h2d(b->cnt + 1, &cnt, 4, st));
    }
#ifndef VIWB_HOST_EMU
    CK(dev_sync(st));                      // `cnt` lives on this stack frame
#endif
    rc = lk_batch_execute(b, mode == 0 ? 1 : 2); if (rc) return rc;
    CK(dev_d2h(pts_b, b->P(mode == 0 ? 1 : 4, 0), (size_t)n * 8, st));
    CK(dev_d2h(status, b->S(mode == 0 ? 0 : 2, 0), (size_t)n, st));
    CK(dev_sync(st));
    return VIWB_OK;
}

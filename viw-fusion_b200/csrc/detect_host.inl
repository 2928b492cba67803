// detect_host.inl -- host side of the feature detector (kernels_detect.cuh), included by viwb.cu after lk_host.inl.
// viwb_detector = device state of F camera streams: mask, candidate list, cell grid, tracked points in,
// surviving indices + new corners out, and the per-stream task table the four kernels index.

#ifdef VIWB_HOST_EMU
static void det_launch_order(const DetRun &r, int F, stream_t) { std::vector<int> order(DET_MAXPTS); for (int f = 0; f < F; f++) det_order_warp(r.tasks[f], r, 0, 1, order.data()); }
static void det_launch_mask(const DetRun &r, int F, int h, stream_t) {
    std::vector<short> nx(2 * DET_MAXPTS); int nn = 0;
    for (int f = 0; f < F; f++) for (int b = 0; b < (h + DET_BAND - 1) / DET_BAND; b++) det_mask_band(r.tasks[f], r, b, 0, 1, nx.data(), &nn);
}
static void det_launch_corners(const DetRun &r, int F, int w, int h, stream_t) {
    std::vector<double> sm(det_tile_smem_bytes() / 8 + 1);
    const int tx = (w + DET_OW - 1) / DET_OW, ty = (h + DET_OH - 1) / DET_OH;
    for (int f = 0; f < F; f++) for (int t = 0; t < tx * ty; t++) det_corner_tile(r.tasks[f], r, t % tx, t / tx, 0, 1, (unsigned char *)sm.data());
}
static void det_launch_select(const DetRun &r, int F, stream_t) {
    std::vector<double> sm(det_select_smem_bytes() / 8 + 1); int counter[2];
    for (int f = 0; f < F; f++) det_select_block(r.tasks[f], r, 0, 1, (unsigned char *)sm.data(), counter);
}
#else
static void det_launch_order(const DetRun &r, int F, stream_t s) { g_prof.begin("det_order", s); det_order_kernel<<<F, 32, 0, s>>>(r); g_prof.end(s); }
static void det_launch_mask(const DetRun &r, int F, int h, stream_t s) { g_prof.begin("det_mask", s); det_mask_kernel<<<dim3((h + DET_BAND - 1) / DET_BAND, F), 256, 0, s>>>(r); g_prof.end(s); }
static void det_launch_corners(const DetRun &r, int F, int w, int h, stream_t s) {
    const int tx = (w + DET_OW - 1) / DET_OW, ty = (h + DET_OH - 1) / DET_OH;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(det_corners_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_tile_smem_bytes()); attr = true; }
    g_prof.begin("det_corners", s); det_corners_kernel<<<dim3(tx * ty, F), DET_NT, det_tile_smem_bytes(), s>>>(r, tx); g_prof.end(s);
}
static void det_launch_select(const DetRun &r, int F, stream_t s) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(det_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_select_smem_bytes()); attr = true; }
    g_prof.begin("det_select", s); det_select_kernel<<<F, 1024, det_select_smem_bytes(), s>>>(r); g_prof.end(s);
}
#endif

// half widths of the rows of cv::circle(..., radius, ..., -1): the midpoint circle of OpenCV's drawing.cpp (third party), restated
static std::vector<short> det_circle_half_widths(int r) {
    std::vector<short> hw((size_t)r + 1, (short)-1);
    int err = 0, dx = r, dy = 0, plus = 1, minus = (r << 1) - 1;
    while (dx >= dy) {
        if (hw[dx] < dy) hw[dx] = (short)dy;
        if (hw[dy] < dx) hw[dy] = (short)dx;
        dy++; err += plus; plus += 2;
        const int m = (err <= 0) - 1;
        err -= minus & m; dx += m; minus -= m & 2;
    }
    return hw;
}

enum { DET_N_PTS = 0, DET_N_KEEP = 1, DET_N_CAND = 2, DET_N_CORNERS = 3, DET_MAXBITS = 4, DET_COUNTERS = 5 };
struct viwb_detector {
    viwb_context *ctx;
    int F, w, h, maxn, radius, cand_cap, gw, gh, cell;
    double min_dist;
    uint8_t *img, *mask, *base;            // [F][h][w] each (img only used when the caller hands host images)
    float *pts, *corners;                  // [F][maxn][2], [F][maxn][2]
    int *track_cnt, *keep, *counters;      // [F][maxn], [F][maxn], [DET_COUNTERS][F]
    short *hw, *kept_xy;                   // [radius+1], [F][maxn][2]
    unsigned long long *cand;              // [F][cand_cap]
    unsigned int *grid;                    // [F][gw*gh][DET_SLOTS]
    DetArgs *tasks;                        // [F]
    const viwb_lk_batch *bound;            // tracker whose resident left images the task table points at (NULL: own `img`)
    size_t bytes;
};

static void det_free(viwb_detector *d) {
    if (!d) return;
    void *p[] = {d->img, d->mask, d->base, d->pts, d->corners, d->track_cnt, d->keep, d->counters, d->hw, d->kept_xy, d->cand, d->grid, d->tasks};
    for (void *q : p) if (q) dev_free(q);
    delete d;
}

static int det_write_tasks(viwb_detector *d, const viwb_lk_batch *lk) {
    viwb_context *ctx = d->ctx;
    std::vector<DetArgs> T((size_t)d->F);
    const size_t px = (size_t)d->w * d->h, cells = (size_t)d->gw * d->gh * DET_SLOTS;
    for (int f = 0; f < d->F; f++) {
        DetArgs a; memset(&a, 0, sizeof a);
        if (lk) { a.img[0] = lk->image(0, 0, f); a.img[1] = lk->image(0, 1, f); a.stride = lk->ls[0]; }
        else { a.img[0] = a.img[1] = d->img + px * f; a.stride = d->w; }
        a.w = d->w; a.h = d->h; a.base_mask = d->base + px * f;
        a.pts = d->pts + (size_t)f * d->maxn * 2; a.track_cnt = d->track_cnt + (size_t)f * d->maxn; a.n_dev = d->counters + (size_t)DET_N_PTS * d->F + f;
        a.radius = d->radius; a.hw = d->hw;
        a.keep = d->keep + (size_t)f * d->maxn; a.n_keep = d->counters + (size_t)DET_N_KEEP * d->F + f; a.kept_xy = d->kept_xy + (size_t)f * d->maxn * 2;
        a.mask = d->mask + px * f; a.maxbits = (unsigned *)(d->counters + (size_t)DET_MAXBITS * d->F + f);
        a.cand = d->cand + (size_t)f * d->cand_cap; a.cand_cap = d->cand_cap; a.n_cand = d->counters + (size_t)DET_N_CAND * d->F + f;
        a.grid = d->grid + cells * f; a.gw = d->gw; a.gh = d->gh; a.cell = d->cell;
        a.corners = d->corners + (size_t)f * d->maxn * 2; a.corner_cap = d->maxn; a.n_corners = d->counters + (size_t)DET_N_CORNERS * d->F + f;
        T[f] = a;
    }
    CK(dev_h2d(d->tasks, T.data(), sizeof(DetArgs) * T.size(), ctx->stream));
    CK(dev_sync(ctx->stream));
    d->bound = lk;
    return VIWB_OK;
}

static int det_build(viwb_context *ctx, int F, int w, int h, int maxn, double min_dist, viwb_detector **out) {
    bind_device(ctx);
    if (F <= 0 || w < 3 || h < 3 || w > 32767 || h > 32767 || maxn <= 0 || maxn > DET_MAXPTS || !(min_dist >= 0.0) || min_dist > 16384.0)
        return fail(ctx, VIWB_ERR_INVALID, "detector: bad geometry (3 <= w,h <= 32767, 1 <= max_pts <= 1024, 0 <= min_dist)");
    viwb_detector *d = new viwb_detector();
    memset(d, 0, sizeof *d);
    d->ctx = ctx; d->F = F; d->w = w; d->h = h; d->maxn = maxn; d->min_dist = min_dist; d->radius = (int)min_dist;
    d->cell = min_dist >= 1.0 ? (int)lrint(min_dist) : 1;                              // cvRound(minDistance)
    d->gw = (w + d->cell - 1) / d->cell; d->gh = (h + d->cell - 1) / d->cell;
    d->cand_cap = det_pow2_at_least((int)(((size_t)w * h + 4) / 5));                   // strict 3x3 maxima cannot be denser than 1 in 4 pixels; 1 in 5 is already pathological
    if (d->cand_cap < 1024) d->cand_cap = 1024;
    const size_t px = (size_t)w * h, cells = (size_t)d->gw * d->gh * DET_SLOTS;
#define DTA(p, n) do { if (dev_malloc((void **)&(p), (n))) { det_free(d); return fail(ctx, VIWB_ERR_CUDA, "detector device allocation failed"); } d->bytes += (n); } while (0)
    DTA(d->img, px * F); DTA(d->mask, px * F); DTA(d->base, px * F);
    DTA(d->pts, (size_t)F * maxn * 8); DTA(d->corners, (size_t)F * maxn * 8); DTA(d->track_cnt, (size_t)F * maxn * 4); DTA(d->keep, (size_t)F * maxn * 4);
    DTA(d->counters, (size_t)DET_COUNTERS * F * 4); DTA(d->hw, (size_t)(d->radius + 1) * 2); DTA(d->kept_xy, (size_t)F * maxn * 4);
    DTA(d->cand, (size_t)F * d->cand_cap * 8); DTA(d->grid, cells * F * 4); DTA(d->tasks, sizeof(DetArgs) * F);
#undef DTA
    const std::vector<short> hw = det_circle_half_widths(d->radius);
    int e = dev_h2d(d->hw, hw.data(), hw.size() * 2, ctx->stream);
    if (!e) e = dev_sync(ctx->stream);
    if (e) { det_free(d); return fail(ctx, VIWB_ERR_CUDA, "detector table upload failed"); }
    const int rc = det_write_tasks(d, nullptr);
    if (rc) { det_free(d); return rc; }
    *out = d;
    return VIWB_OK;
}

static int det_upload_images(viwb_detector *d, uint8_t *dst, const uint8_t *const *imgs, int stride) {
    viwb_context *ctx = d->ctx;
    const size_t px = (size_t)d->w * d->h;
    bool strided = stride == d->w;
    for (int f = 1; f < d->F && strided; f++) strided = imgs[f] == imgs[0] + px * f;
    if (strided) { CK(dev_h2d(dst, imgs[0], px * d->F, ctx->stream)); return VIWB_OK; }
    for (int f = 0; f < d->F; f++) CK(dev_h2d_2d(dst + px * f, d->w, imgs[f], stride, d->w, d->h, ctx->stream));
    return VIWB_OK;
}

// the four launches of one tick over inputs that already sit in the detector's device buffers (points, track counts, counts)
static void det_run_device(viwb_detector *d, const viwb_lk_batch *resident, int max_cnt, double quality, int use_base) {
    viwb_context *ctx = d->ctx; stream_t st = ctx->stream;
    DetRun run; memset(&run, 0, sizeof run);
    run.tasks = d->tasks; run.img_sel = resident ? resident->cur : 0; run.use_mask = 1; run.use_base = use_base; run.tracker_mode = 1;
    run.max_cnt = max_cnt; run.quality = quality; run.min_dist = d->min_dist;
    det_launch_order(run, d->F, st); det_launch_mask(run, d->F, d->h, st); det_launch_corners(run, d->F, d->w, d->h, st); det_launch_select(run, d->F, st);
    ctx->launches += 4;
}

// One camera tick of F streams: setMask over the tracked points, goodFeaturesToTrack for the MAX_CNT - n_keep missing corners.
static int det_detect(viwb_detector *d, const uint8_t *const *images, int stride, const viwb_lk_batch *resident, const uint8_t *const *base_masks,
                      const float *pts, const int *track_cnt, const int *n_pts, int max_cnt, double quality,
                      int *keep, int *n_keep, float *new_pts, int *n_new, uint8_t *mask_out) {
    viwb_context *ctx = d->ctx;
    bind_device(ctx);
    if (!images && !resident) return fail(ctx, VIWB_ERR_INVALID, "detector: neither host images nor a resident tracker given");
    if (resident && (resident->w != d->w || resident->h != d->h || resident->F != d->F)) return fail(ctx, VIWB_ERR_INVALID, "detector: the tracker's geometry differs");
    for (int f = 0; f < d->F; f++) if (n_pts[f] < 0 || n_pts[f] > d->maxn) return fail(ctx, VIWB_ERR_INVALID, "detector: point count exceeds max_pts");
    if (max_cnt > d->maxn) return fail(ctx, VIWB_ERR_INVALID, "detector: max_cnt exceeds max_pts");
    const viwb_lk_batch *want = images ? nullptr : resident;
    if (d->bound != want) { const int rc = det_write_tasks(d, want); if (rc) return rc; }
    stream_t st = ctx->stream;
    if (images) { const int rc = det_upload_images(d, d->img, images, stride); if (rc) return rc; }
    if (base_masks) { const int rc = det_upload_images(d, d->base, base_masks, d->w); if (rc) return rc; }
    const size_t np = (size_t)d->F * d->maxn;
    CK(dev_h2d(d->pts, pts, np * 8, st)); CK(dev_h2d(d->track_cnt, track_cnt, np * 4, st)); CK(dev_h2d(d->counters + (size_t)DET_N_PTS * d->F, n_pts, (size_t)d->F * 4, st));
    det_run_device(d, want, max_cnt, quality, base_masks ? 1 : 0);
    CK(dev_d2h(keep, d->keep, np * 4, st)); CK(dev_d2h(n_keep, d->counters + (size_t)DET_N_KEEP * d->F, (size_t)d->F * 4, st));
    CK(dev_d2h(new_pts, d->corners, np * 8, st)); CK(dev_d2h(n_new, d->counters + (size_t)DET_N_CORNERS * d->F, (size_t)d->F * 4, st));
    if (mask_out) CK(dev_d2h(mask_out, d->mask, (size_t)d->w * d->h * d->F, st));
    CK(dev_sync(st));
    for (int f = 0; f < d->F; f++) if (n_new[f] < 0) return fail(ctx, VIWB_ERR_INVALID, "detector: candidate list overflow (image denser in local maxima than 1 per 5 pixels)");
    return VIWB_OK;
}

// ---- single-image entry points with the cv:: signatures; a one-stream detector is cached in the context
static int det_single(viwb_context *ctx, int w, int h, int n, double min_dist, viwb_detector **out) {
    viwb_detector *d = ctx->det1;
    if (d && (d->w != w || d->h != h || d->maxn < n || d->min_dist != min_dist)) { det_free(d); ctx->det1 = d = nullptr; }
    if (!d) {
        int cap = 256; while (cap < n) cap *= 2;
        if (cap > DET_MAXPTS) cap = DET_MAXPTS;
        const int rc = det_build(ctx, 1, w, h, cap, min_dist, &d); if (rc) return rc;
        ctx->det1 = d;
    }
    if (d->bound) { const int rc = det_write_tasks(d, nullptr); if (rc) return rc; }
    *out = d;
    return VIWB_OK;
}

static int det_good_features(viwb_context *ctx, const uint8_t *img, int w, int h, int stride, int max_corners, double quality, double min_dist,
                             const uint8_t *mask, int mask_stride, float *corners, int capacity, int *n_corners) {
    bind_device(ctx);
    if (!img || !corners || !n_corners || capacity <= 0 || capacity > DET_MAXPTS || quality <= 0.0 || min_dist < 0.0) return fail(ctx, VIWB_ERR_INVALID, "goodFeaturesToTrack: bad argument (1 <= capacity <= 1024, quality > 0, min_dist >= 0)");
    viwb_detector *d; int rc = det_single(ctx, w, h, capacity, min_dist, &d); if (rc) return rc;
    stream_t st = ctx->stream;
    CK(dev_h2d_2d(d->img, w, img, stride, w, h, st));
    if (mask) CK(dev_h2d_2d(d->mask, w, mask, mask_stride, w, h, st));
    const int zero[DET_COUNTERS] = {0, 0, 0, 0, 0};
    CK(dev_h2d(d->counters, zero, sizeof zero, st));
    DetRun run; memset(&run, 0, sizeof run);
    run.tasks = d->tasks; run.use_mask = mask ? 1 : 0; run.max_corners = max_corners; run.corner_cap = capacity; run.quality = quality; run.min_dist = min_dist;
    det_launch_corners(run, 1, w, h, st); det_launch_select(run, 1, st);
    ctx->launches += 2;
    int cnt = 0;
    CK(dev_d2h(&cnt, d->counters + DET_N_CORNERS, 4, st)); CK(dev_sync(st));
    if (cnt < 0) return fail(ctx, VIWB_ERR_INVALID, "goodFeaturesToTrack: candidate list overflow");
    CK(dev_d2h(corners, d->corners, (size_t)cnt * 8, st)); CK(dev_sync(st));
    *n_corners = cnt;
    return VIWB_OK;
}

static int det_set_mask(viwb_context *ctx, int w, int h, const float *pts, const int *track_cnt, int n, int min_dist, const uint8_t *base_mask,
                        uint8_t *mask_out, int *keep, int *n_keep) {
    bind_device(ctx);
    if (n < 0 || n > DET_MAXPTS || min_dist < 0 || !keep || !n_keep || (n && (!pts || !track_cnt))) return fail(ctx, VIWB_ERR_INVALID, "setMask: bad argument (0 <= n <= 1024)");
    viwb_detector *d; int rc = det_single(ctx, w, h, n, (double)min_dist, &d); if (rc) return rc;
    stream_t st = ctx->stream;
    const int cnts[DET_COUNTERS] = {n, 0, 0, 0, 0};
    CK(dev_h2d(d->counters, cnts, sizeof cnts, st));
    CK(dev_h2d(d->pts, pts, (size_t)n * 8, st)); CK(dev_h2d(d->track_cnt, track_cnt, (size_t)n * 4, st));
    if (base_mask) CK(dev_h2d(d->base, base_mask, (size_t)w * h, st));
    DetRun run; memset(&run, 0, sizeof run);
    run.tasks = d->tasks; run.use_base = base_mask ? 1 : 0; run.min_dist = (double)min_dist;
    det_launch_order(run, 1, st); det_launch_mask(run, 1, h, st);
    ctx->launches += 2;
    int nk = 0;
    CK(dev_d2h(&nk, d->counters + DET_N_KEEP, 4, st)); CK(dev_sync(st));
    CK(dev_d2h(keep, d->keep, (size_t)nk * 4, st));
    if (mask_out) CK(dev_d2h(mask_out, d->mask, (size_t)w * h, st));
    CK(dev_sync(st));
    *n_keep = nk;
    return VIWB_OK;
}

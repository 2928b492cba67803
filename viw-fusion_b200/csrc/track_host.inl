// track_host.inl -- host side of the session tracker (kernels_track.cuh), included by viwb.cu after detect_host.inl.
// viwb_tracker = F independent FeatureTracker sessions (featureTracker/feature_tracker.h): a viwb_lk_batch (images, pyramids, flow
// buffers), a viwb_detector bound to the tracker's resident images, and per-stream session state in HBM (ids, track counts, n_id,
// the id -> undistorted point maps of the previous tick).  One viwb_tracker_track() = one FeatureTracker::trackImage() per stream:
// images in, featureFrame rows out, nothing else crosses the bus.

#ifdef VIWB_HOST_EMU
static void trk_launch_pred_setup(const TrkRun &r, int F, stream_t) { for (int f = 0; f < F; f++) trk_pred_setup(r.tasks[f], 0); }
static void trk_launch_pred_check(const TrkRun &r, int F, stream_t) { for (int f = 0; f < F; f++) trk_pred_check(r.tasks[f], 0, 1); }
static void trk_launch_advance(const TrkRun &r, int F, stream_t) { for (int f = 0; f < F; f++) trk_advance(r.tasks[f], 0, 1); }
static void trk_launch_merge(const TrkRun &r, int F, stream_t) { for (int f = 0; f < F; f++) trk_merge(r.tasks[f], r, 0, 1); }
static void trk_launch_stereo(const TrkRun &r, int F, stream_t) { for (int f = 0; f < F; f++) trk_stereo(r.tasks[f], r, 0, 1); }
#else
static void trk_launch_pred_setup(const TrkRun &r, int F, stream_t s) { g_prof.begin("trk_pred_setup", s); trk_pred_setup_kernel<<<F, 32, 0, s>>>(r); g_prof.end(s); }
static void trk_launch_pred_check(const TrkRun &r, int F, stream_t s) { g_prof.begin("trk_pred_check", s); trk_pred_check_kernel<<<F, 32, 0, s>>>(r); g_prof.end(s); }
static void trk_launch_advance(const TrkRun &r, int F, stream_t s) { g_prof.begin("trk_advance", s); trk_advance_kernel<<<F, 32, 0, s>>>(r); g_prof.end(s); }
static void trk_launch_merge(const TrkRun &r, int F, stream_t s) { g_prof.begin("trk_merge", s); trk_merge_kernel<<<F, 128, 0, s>>>(r); g_prof.end(s); }
static void trk_launch_stereo(const TrkRun &r, int F, stream_t s) { g_prof.begin("trk_stereo", s); trk_stereo_kernel<<<F, 32, 0, s>>>(r); g_prof.end(s); }
#endif

enum { TRK_N_ID = 0, TRK_MAPL_N, TRK_MAPR_N, TRK_PRED_N, TRK_REDO_N, TRK_OUT_N, TRK_OUT_NR, TRK_FLAGS, TRK_COUNTERS };
struct viwb_tracker {
    viwb_context *ctx;
    int F, w, h, maxn, max_cnt, stereo, flow_back;
    double quality, prev_time; bool has_prev_time;
    TrkCam cam[2];
    viwb_lk_batch *lk; viwb_detector *det;
    int *ints;                 // [6][F][maxn]: ids, track_cnt, tmp_ids, mapL_ids, mapR_ids, out_ids_r  (out_ids / out_cnt follow)
    int *outs;                 // [2][F][maxn]: out_ids, out_cnt
    float *mapun;              // [2][F][maxn][2]
    float *feat;               // [2][F][maxn][6]
    int *counters;             // [TRK_COUNTERS][F]
    uint8_t *has_pred;         // [F]
    TrkArgs *tasks;            // [F]
    LkArgs *pred_tasks;        // [2 (cur)][2 (seeded pass / full-pyramid repeat)][F]
    long long ticks;
};

static void trk_free(viwb_tracker *t) {
    if (!t) return;
    void *p[] = {t->ints, t->outs, t->mapun, t->feat, t->counters, t->has_pred, t->tasks, t->pred_tasks};
    for (void *q : p) if (q) dev_free(q);
    if (t->det) det_free(t->det);
    if (t->lk) lk_batch_free(t->lk);
    delete t;
}

static int trk_build(viwb_context *ctx, int F, int w, int h, const viwb_tracker_config *cfg, viwb_tracker **out) {
    bind_device(ctx);
    if (!cfg || F <= 0 || cfg->max_cnt <= 0 || cfg->max_cnt > DET_MAXPTS || cfg->min_dist < 0) return fail(ctx, VIWB_ERR_INVALID, "tracker: bad configuration (1 <= max_cnt <= 1024, min_dist >= 0)");
    viwb_tracker *t = new viwb_tracker();
    memset(t, 0, sizeof *t);
    t->ctx = ctx; t->F = F; t->w = w; t->h = h; t->maxn = cfg->max_cnt; t->max_cnt = cfg->max_cnt; t->stereo = cfg->stereo ? 1 : 0; t->flow_back = cfg->flow_back ? 1 : 0;
    t->quality = 0.01;                                                  // feature_tracker.cpp:192
    for (int c = 0; c < 2; c++) { const viwb_pinhole &p = cfg->cam[c]; TrkCam k = {p.fx, p.fy, p.cx, p.cy, p.k1, p.k2, p.p1, p.p2}; t->cam[c] = k; }
    int rc = lk_batch_build(ctx, F, w, h, t->maxn, t->stereo, t->flow_back, &t->lk); if (rc) { trk_free(t); return rc; }
    rc = det_build(ctx, F, w, h, t->maxn, (double)cfg->min_dist, &t->det); if (rc) { trk_free(t); return rc; }
    rc = det_write_tasks(t->det, t->lk); if (rc) { trk_free(t); return rc; }
    const size_t np = (size_t)F * t->maxn;
#define TKA(p, n) do { if (dev_malloc((void **)&(p), (n))) { trk_free(t); return fail(ctx, VIWB_ERR_CUDA, "tracker device allocation failed"); } } while (0)
    TKA(t->ints, np * 6 * 4); TKA(t->outs, np * 2 * 4); TKA(t->mapun, np * 2 * 8); TKA(t->feat, np * 2 * 24); TKA(t->counters, (size_t)TRK_COUNTERS * F * 4);
    TKA(t->has_pred, (size_t)F); TKA(t->tasks, sizeof(TrkArgs) * F); TKA(t->pred_tasks, sizeof(LkArgs) * 4 * F);
#undef TKA
    viwb_lk_batch *b = t->lk; viwb_detector *d = t->det;
    std::vector<TrkArgs> T((size_t)F);
    for (int f = 0; f < F; f++) {
        TrkArgs a; memset(&a, 0, sizeof a);
        a.prev_pts = b->P(0, f); a.lk_cur = b->P(1, f); a.lk_status = b->S(0, f); a.n_temporal = b->cnt + f;
        a.stereo_in = b->P(3, f); a.right_pts = b->P(4, f); a.st_right = b->S(2, f); a.n_stereo = b->cnt + F + f;
        a.det_pts = d->pts + (size_t)f * d->maxn * 2; a.det_cnt = d->track_cnt + (size_t)f * d->maxn; a.det_n = d->counters + (size_t)DET_N_PTS * F + f;
        a.keep = d->keep + (size_t)f * d->maxn; a.n_keep = d->counters + (size_t)DET_N_KEEP * F + f;
        a.corners = d->corners + (size_t)f * d->maxn * 2; a.n_corners = d->counters + (size_t)DET_N_CORNERS * F + f;
        int *I = t->ints + (size_t)f * t->maxn;
        a.ids = I; a.track_cnt = I + np; a.tmp_ids = I + 2 * np; a.mapL_ids = I + 3 * np; a.mapR_ids = I + 4 * np; a.out_ids_r = I + 5 * np;
        a.out_ids = t->outs + (size_t)f * t->maxn; a.out_cnt = a.out_ids + np;
        a.mapL_un = t->mapun + (size_t)f * t->maxn * 2; a.mapR_un = a.mapL_un + np * 2;
        a.out_feat = t->feat + (size_t)f * t->maxn * 6; a.out_feat_r = a.out_feat + np * 6;
        int *Cn = t->counters + f;
        a.n_id = Cn + (size_t)TRK_N_ID * F; a.mapL_n = Cn + (size_t)TRK_MAPL_N * F; a.mapR_n = Cn + (size_t)TRK_MAPR_N * F; a.pred_n = Cn + (size_t)TRK_PRED_N * F;
        a.redo_n = Cn + (size_t)TRK_REDO_N * F; a.out_n = Cn + (size_t)TRK_OUT_N * F; a.out_nr = Cn + (size_t)TRK_OUT_NR * F; a.flags = Cn + (size_t)TRK_FLAGS * F;
        a.has_pred = t->has_pred + f;
        T[f] = a;
    }
    // the two passes of the hasPrediction branch (feature_tracker.cpp:122-137): same buffers as the tracker's temporal forward task, other counts
    std::vector<LkArgs> L((size_t)4 * F), W((size_t)8 * F);
    int e = dev_d2h(W.data(), b->tasks, sizeof(LkArgs) * 8 * F, ctx->stream);
    if (!e) e = dev_sync(ctx->stream);
    for (int cur = 0; cur < 2 && !e; cur++) for (int f = 0; f < F; f++) {
        LkArgs a = W[((size_t)cur * 2 + 0) * 2 * F + f];                 // temporal forward of this parity
        LkArgs p = a; p.n_dev = T[f].pred_n; p.max_level = b->levels < 1 ? b->levels : 1; p.flags = 4;
        LkArgs r = a; r.n_dev = T[f].redo_n;
        L[((size_t)cur * 2 + 0) * F + f] = p; L[((size_t)cur * 2 + 1) * F + f] = r;
    }
    const std::vector<int> zero((size_t)TRK_COUNTERS * F, 0); const std::vector<int> zero2((size_t)2 * F, 0);
    if (!e) e = dev_h2d(t->tasks, T.data(), sizeof(TrkArgs) * F, ctx->stream);
    if (!e) e = dev_h2d(t->pred_tasks, L.data(), sizeof(LkArgs) * 4 * F, ctx->stream);
    if (!e) e = dev_h2d(t->counters, zero.data(), zero.size() * 4, ctx->stream);
    if (!e) e = dev_h2d(b->cnt, zero2.data(), zero2.size() * 4, ctx->stream);
    if (!e) e = dev_sync(ctx->stream);
    if (e) { trk_free(t); return fail(ctx, VIWB_ERR_CUDA, "tracker table upload failed"); }
    *out = t;
    return VIWB_OK;
}

// One FeatureTracker::trackImage(cur_time, left, right) for every stream; asynchronous on the context stream until the download
static int trk_track(viwb_tracker *t, double cur_time, const uint8_t *const *left, const uint8_t *const *right, int stride, const float *predict_pts,
                     const uint8_t *has_prediction) {
    viwb_context *ctx = t->ctx;
    bind_device(ctx);
    viwb_lk_batch *b = t->lk; const int F = t->F; stream_t st = ctx->stream;
    if (!left) return fail(ctx, VIWB_ERR_INVALID, "tracker: left images missing");
    if (t->stereo && !right) return fail(ctx, VIWB_ERR_INVALID, "tracker: stereo session without right images");
    int rc = lk_batch_upload(b, nullptr, left, t->stereo ? right : nullptr, stride, nullptr, nullptr, nullptr, nullptr); if (rc) return rc;
    TrkRun run; memset(&run, 0, sizeof run);
    run.tasks = t->tasks; run.cam[0] = t->cam[0]; run.cam[1] = t->cam[1]; run.maxn = t->maxn;
    run.dt = t->has_prev_time ? cur_time - t->prev_time : 1.0;
    if (t->ticks > 0) {
        bool any_pred = false;
        if (predict_pts && has_prediction) for (int f = 0; f < F; f++) any_pred = any_pred || has_prediction[f];
        if (!any_pred) { rc = lk_batch_execute(b, 1, false); if (rc) return rc; }
        else {
            // pyramids of the new images, then: seeded maxLevel-1 pass for the streams with a prediction, full-pyramid pass for the others
            // and for those with fewer than 10 successes, reverse flow and status rules for everyone
            for (int s = 0; s < 2; s++) if (b->dirty[s]) {
                for (int l = 1; l <= b->levels; l++) { lk_launch_pyr(b->pyr + ((size_t)s * 3 + (l - 1)) * F, pyr_items_wh(b->lw[l], b->lh[l]), F, st); ctx->launches++; }
                b->dirty[s] = false;
            }
            CK(dev_h2d(t->has_pred, has_prediction, (size_t)F, st));
            CK(dev_h2d(b->P(1, 0), predict_pts, (size_t)F * t->maxn * 8, st));
            const LkArgs *pt = t->pred_tasks + ((size_t)b->cur * 2 + 0) * F, *rt = t->pred_tasks + ((size_t)b->cur * 2 + 1) * F;
            trk_launch_pred_setup(run, F, st); lk_launch_track(pt, t->maxn, F, st, b->maps, b->use_tma);
            trk_launch_pred_check(run, F, st); lk_launch_track(rt, t->maxn, F, st, b->maps, b->use_tma);
            ctx->launches += 4;
            const LkArgs *w2 = b->tasks + ((size_t)b->cur * 2 + 1) * 2 * F;
            if (b->flow_back) { CK(dev_d2d(b->P(2, 0), b->P(0, 0), (size_t)F * b->maxn * 8, st)); lk_launch_track(w2, t->maxn, F, st, b->maps, b->use_tma); ctx->launches++; }
            lk_launch_post(b->post, t->maxn, F, st); ctx->launches++;
        }
    }
    trk_launch_advance(run, F, st); ctx->launches++;
    det_run_device(t->det, b, t->max_cnt, t->quality, 0);
    trk_launch_merge(run, F, st); ctx->launches++;
    if (t->stereo) {
        rc = lk_batch_execute(b, 2, false); if (rc) return rc;
        trk_launch_stereo(run, F, st); ctx->launches++;
    }
#ifndef VIWB_HOST_EMU
    CK((int)cudaGetLastError());
#endif
    t->prev_time = cur_time; t->has_prev_time = true; t->ticks++;
    return VIWB_OK;
}

static int trk_fetch(viwb_tracker *t, int32_t *n_left, int32_t *ids, int32_t *track_cnt, float *feat, int32_t *n_right, int32_t *ids_right, float *feat_right) {
    viwb_context *ctx = t->ctx;
    bind_device(ctx);
    const int F = t->F; const size_t np = (size_t)F * t->maxn; stream_t st = ctx->stream;
    std::vector<int> flags((size_t)F, 0);
    if (n_left) CK(dev_d2h(n_left, t->counters + (size_t)TRK_OUT_N * F, (size_t)F * 4, st));
    if (ids) CK(dev_d2h(ids, t->outs, np * 4, st));
    if (track_cnt) CK(dev_d2h(track_cnt, t->outs + np, np * 4, st));
    if (feat) CK(dev_d2h(feat, t->feat, np * 24, st));
    if (n_right) CK(dev_d2h(n_right, t->counters + (size_t)TRK_OUT_NR * F, (size_t)F * 4, st));
    if (ids_right) CK(dev_d2h(ids_right, t->ints + 5 * np, np * 4, st));
    if (feat_right) CK(dev_d2h(feat_right, t->feat + np * 6, np * 24, st));
    CK(dev_d2h(flags.data(), t->counters + (size_t)TRK_FLAGS * F, (size_t)F * 4, st));
    CK(dev_sync(st));
    for (int f = 0; f < F; f++) if (flags[f] & 1) return fail(ctx, VIWB_ERR_INVALID, "tracker: goodFeaturesToTrack candidate list overflow (image denser in local maxima than 1 per 5 pixels)");
    return VIWB_OK;
}

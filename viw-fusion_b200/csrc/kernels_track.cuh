// kernels_track.cuh -- the bookkeeping of FeatureTracker::trackImage() (featureTracker/feature_tracker.cpp:99-331) that sits between
// the calcOpticalFlowPyrLK / setMask / goodFeaturesToTrack calls, so that a camera tick of every stream runs on the device from the
// uploaded images to the featureFrame table without a host round trip:
//   reduceVector(prev_pts / cur_pts / ids / track_cnt, status) + track_cnt++                      :163-171
//   rebuild of cur_pts / ids / track_cnt in setMask's visiting order + the new corners (ids n_id++) :72-88, :199-204
//   undistortedPts + ptsVelocity through the id -> point maps of the previous tick                 :207-208, :606-657
//   the stereo half: ids_right = ids, reduceVector by the stereo status, right undistort / velocity  :254-268
//   prev_* = cur_* (:296-300) and the featureFrame rows (:307-350)
// hasPrediction (:122-137): forward flow seeded with predict_pts at maxLevel 1, repeated at maxLevel 3 when fewer than 10 succeed.
//
// One block per camera stream (these are 150-point lists); `tid` / `nt` = thread index / count, the emulation build runs nt = 1.
#pragma once
#include <stdint.h>
#include "vmath.cuh"
#include "kernels_feat.cuh"

namespace viwb {

struct TrkCam { double fx, fy, cx, cy, k1, k2, p1, p2; };

struct TrkArgs {                         // one camera stream; every array holds `maxn` entries
    // --- tracker buffers (viwb_lk_batch): prev_pts = points of the previous image, lk_cur = temporal LK result, stereo_in / right = stereo LK in / out
    float *prev_pts; const float *lk_cur; const uint8_t *lk_status; int *n_temporal;
    float *stereo_in; const float *right_pts; const uint8_t *st_right; int *n_stereo;
    // --- detector buffers (viwb_detector): setMask input, survivors in visiting order, new corners
    float *det_pts; int *det_cnt; int *det_n; const int *keep; const int *n_keep; const float *corners; const int *n_corners;
    // --- session state
    int *ids, *track_cnt, *tmp_ids;      // aligned with prev_pts; tmp_ids aligned with det_pts
    int *n_id;                           // next feature id of this stream (feature_tracker.h: n_id)
    int *mapL_ids; float *mapL_un; int *mapL_n;      // prev_un_pts_map
    int *mapR_ids; float *mapR_un; int *mapR_n;      // prev_un_right_pts_map
    // --- prediction
    const uint8_t *has_pred; int *pred_n, *redo_n;
    // --- featureFrame rows: {x, y, u, v, vx, vy} (z = 1), left rows in cur_pts order, right rows in ids_right order
    int *out_n; int *out_ids, *out_cnt; float *out_feat;
    int *out_nr; int *out_ids_r; float *out_feat_r;
    int *flags;                          // bit 0: goodFeaturesToTrack candidate list overflowed in this tick
};
struct TrkRun { const TrkArgs *tasks; TrkCam cam[2]; double dt; int maxn; };

// warp-cooperative order-preserving position of an element with `flag` among the flagged ones of a group of nt consecutive indices
VIWB_D int trk_rank(bool flag, int tid, int &total) {
#ifdef VIWB_HOST_EMU
    (void)tid; total = flag ? 1 : 0; return 0;
#else
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    total = __popc(m);
    return __popc(m & ((1u << tid) - 1u));
#endif
}

// PinholeCamera::liftProjective (camera_models/src/camera_models/PinholeCamera.cc:450-517) narrowed to cv::Point2f (feature_tracker.cpp:614)
VIWB_D void trk_lift(const TrkCam &c, float u, float v, float &x, float &y) {
    UndistArgs a; a.n = 1; a.pts = nullptr; a.prev_un = nullptr; a.has_prev = nullptr; a.vel = nullptr;
    a.fx = c.fx; a.fy = c.fy; a.cx = c.cx; a.cy = c.cy; a.k1 = c.k1; a.k2 = c.k2; a.p1 = c.p1; a.p2 = c.p2; a.dt = 1.0;
    const float p[2] = {u, v}; float un[2];
    a.pts = p; a.un = un;
    undistort_item(a, 0);
    x = un[0]; y = un[1];
}

// ptsVelocity (:619-657): (cur - prev) / dt for ids found in the previous tick's map, float difference and double division, else 0
VIWB_D void trk_velocity(const int *map_ids, const float *map_un, int map_n, int id, float ux, float uy, double dt, float &vx, float &vy) {
    vx = 0.f; vy = 0.f;
    for (int k = 0; k < map_n; k++) if (map_ids[k] == id) {
        vx = (float)((double)(ux - map_un[2 * k]) / dt);
        vy = (float)((double)(uy - map_un[2 * k + 1]) / dt);
        return;
    }
}

// ---- stage 0 (only when a prediction was set): point counts of the seeded maxLevel-1 pass
VIWB_D void trk_pred_setup(const TrkArgs &a, int tid) {
    if (tid != 0) return;
    const int n = *a.n_temporal;
    const bool hp = a.has_pred && *a.has_pred;
    *a.pred_n = hp ? n : 0;
    *a.redo_n = hp ? 0 : n;
}
// succ_num < 10 -> the stream repeats the forward flow over the full pyramid (:130-137)
VIWB_D void trk_pred_check(const TrkArgs &a, int tid, int nt) {
    const int n = *a.pred_n;
    if (n == 0) return;                                       // stream without a prediction: redo_n already holds its count
    int succ = 0;
    for (int base = 0; base < n; base += nt) {
        const int i = base + tid; int tot;
        trk_rank(i < n && a.lk_status[i] != 0, tid, tot);
        succ += tot;
    }
    if (tid == 0) *a.redo_n = succ < 10 ? n : 0;
}

// ---- stage 1, after the temporal flow: reduceVector by status and track_cnt++ straight into the detector's setMask input
VIWB_D void trk_advance(const TrkArgs &a, int tid, int nt) {
    const int n = *a.n_temporal;
    int m = 0;
    for (int base = 0; base < n; base += nt) {
        const int i = base + tid; int tot;
        const bool ok = i < n && a.lk_status[i] != 0;
        const int j = m + trk_rank(ok, tid, tot);
        if (ok) {
            a.det_pts[2 * j] = a.lk_cur[2 * i]; a.det_pts[2 * j + 1] = a.lk_cur[2 * i + 1];
            a.det_cnt[j] = a.track_cnt[i] + 1;
            a.tmp_ids[j] = a.ids[i];
        }
        m += tot;
    }
    if (tid == 0) { *a.det_n = m; *a.flags = 0; }
}

// ---- stage 2, after setMask + goodFeaturesToTrack: the tick's cur_pts / ids / track_cnt, undistortedPts, ptsVelocity, left featureFrame
// rows, the stereo flow's input and next tick's prev_pts.  Needs a block barrier between the map lookups and the map replacement.
VIWB_D void trk_merge(const TrkArgs &a, const TrkRun &r, int tid, int nt) {
    const int nk = *a.n_keep;
    int nc = *a.n_corners;
    if (nc < 0) { nc = 0; if (tid == 0) *a.flags |= 1; }
    int total = nk + nc;
    if (total > r.maxn) { total = r.maxn; nc = total - nk; }
    const int id0 = *a.n_id, map_n = *a.mapL_n;
    for (int j = tid; j < total; j += nt) {
        float u, v; int id, cnt;
        if (j < nk) { const int s = a.keep[j]; u = a.det_pts[2 * s]; v = a.det_pts[2 * s + 1]; id = a.tmp_ids[s]; cnt = a.det_cnt[s]; }
        else { u = a.corners[2 * (j - nk)]; v = a.corners[2 * (j - nk) + 1]; id = id0 + (j - nk); cnt = 1; }
        a.prev_pts[2 * j] = u; a.prev_pts[2 * j + 1] = v;
        a.stereo_in[2 * j] = u; a.stereo_in[2 * j + 1] = v;
        a.ids[j] = id; a.track_cnt[j] = cnt;
        float ux, uy, vx = 0.f, vy = 0.f;
        trk_lift(r.cam[0], u, v, ux, uy);
        if (j < nk) trk_velocity(a.mapL_ids, a.mapL_un, map_n, id, ux, uy, r.dt, vx, vy);     // a new corner's id is in no map
        a.out_ids[j] = id; a.out_cnt[j] = cnt;
        float *o = a.out_feat + 6 * j;
        o[0] = ux; o[1] = uy; o[2] = u; o[3] = v; o[4] = vx; o[5] = vy;
    }
    VIWB_SYNC();
    for (int j = tid; j < total; j += nt) { a.mapL_ids[j] = a.out_ids[j]; a.mapL_un[2 * j] = a.out_feat[6 * j]; a.mapL_un[2 * j + 1] = a.out_feat[6 * j + 1]; }
    if (tid == 0) { *a.n_id = id0 + nc; *a.mapL_n = total; *a.out_n = total; *a.n_temporal = total; *a.n_stereo = total; *a.out_nr = 0; }
}

// ---- stage 3, after the stereo flow: ids_right / cur_right_pts by status, right undistort + velocity, right featureFrame rows
VIWB_D void trk_stereo(const TrkArgs &a, const TrkRun &r, int tid, int nt) {
    const int n = *a.n_stereo, map_n = *a.mapR_n;
    int m = 0;
    for (int base = 0; base < n; base += nt) {
        const int i = base + tid; int tot;
        const bool ok = i < n && a.st_right[i] != 0;
        const int j = m + trk_rank(ok, tid, tot);
        if (ok) {
            const float u = a.right_pts[2 * i], v = a.right_pts[2 * i + 1];
            const int id = a.ids[i];
            float ux, uy, vx, vy;
            trk_lift(r.cam[1], u, v, ux, uy);
            trk_velocity(a.mapR_ids, a.mapR_un, map_n, id, ux, uy, r.dt, vx, vy);
            a.out_ids_r[j] = id;
            float *o = a.out_feat_r + 6 * j;
            o[0] = ux; o[1] = uy; o[2] = u; o[3] = v; o[4] = vx; o[5] = vy;
        }
        m += tot;
    }
    VIWB_SYNC();
    for (int j = tid; j < m; j += nt) { a.mapR_ids[j] = a.out_ids_r[j]; a.mapR_un[2 * j] = a.out_feat_r[6 * j]; a.mapR_un[2 * j + 1] = a.out_feat_r[6 * j + 1]; }
    if (tid == 0) { *a.mapR_n = m; *a.out_nr = m; }
}

#ifndef VIWB_HOST_EMU
// one warp per stream for the compactions (ballot ranks), one 128-thread block per stream for the merge
__global__ void trk_pred_setup_kernel(TrkRun r) { trk_pred_setup(r.tasks[blockIdx.x], threadIdx.x); }
__global__ void trk_pred_check_kernel(TrkRun r) { trk_pred_check(r.tasks[blockIdx.x], threadIdx.x, 32); }
__global__ void trk_advance_kernel(TrkRun r) { trk_advance(r.tasks[blockIdx.x], threadIdx.x, 32); }
__global__ void trk_merge_kernel(TrkRun r) { trk_merge(r.tasks[blockIdx.x], r, threadIdx.x, blockDim.x); }
__global__ void trk_stereo_kernel(TrkRun r) { trk_stereo(r.tasks[blockIdx.x], r, threadIdx.x, 32); }
#endif

}  // namespace viwb

#!/usr/bin/env python
"""bench.py -- sliding-window solves/sec (10 key-frames, ~150 features/frame) on N B200s.

A *step* is one pass of the hot path over one batch: B independent sequences each run one
Estimator::optimization() (8 dogleg iterations + gauge re-anchor + MARGIN_OLD marginalisation) on an
EuRoC-shaped stereo+IMU window (BASELINE.json configs[1]) that already carries a marginalisation prior
(steady state), plus -- unless --no-lk -- the frame's four LK calls on 752x480 images.
    value : whole-job solves/s with the batch resident in HBM (CUDA events on the launching stream, max over ranks)
    e2e   : the same through the host-buffer C-ABI call viwb_optimization_batch (lowering + H2D + solve + D2H inside)
    --impl reference : the CPU arm = the oracle (C FP64 port of the reference algorithm) on all host cores.
Inputs: seeded synthetic sequences, G distinct sequences x R perturbed initial guesses each; the
per-step working set (several MB per window) is far larger than L2 at the default batch, which is what keeps
successive steps from hitting in L2 (no explicit flush).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "viw-fusion_b200", "python"))

from viwb import abi, synth  # noqa: E402

METRIC = "sliding-window solves/sec (10 KF, 150 feat)"
UNIT = "solves/s"
OPT = "Estimator::optimization() on a window with prior (8 dogleg iterations + gauge re-anchor + MARGIN_OLD marginalisation)"
# BASELINE.json configs[0..4] = C1..C5 (+ C6: the stereo-only USE_IMU = 0 window, a parity case with a throughput line).  The metric is quoted on C2,
# which is therefore the default at every N; the other configs are selected with --config and reported one line each (profiles/).
WORKLOADS = {
    1: "C1 single 10-keyframe window shape, mono+IMU, 100 landmarks seen in every frame (the reference's own CPU-runnable case), batched: " + OPT + "; no camera tick",
    2: "C2 EuRoC-shaped stereo+IMU stream, 150 feat/frame: per window one camera tick of LK (4 calcOpticalFlowPyrLK-equivalent passes on 752x480 stereo frames) + " + OPT,
    3: "C3 mono+IMU+wheel (realsense_d435i shapes), camera extrinsic + td estimated online: per window one camera tick of LK (2 passes on 640x480 frames) + " + OPT,
    4: "C4 stereo+IMU+wheel with PlaneFactor and MarginalizationFactor active: per window one camera tick of LK (4 passes on 640x480 stereo frames) + " + OPT,
    5: "C5 independent C4-shaped sequences sharded one-per-GPU: per window one camera tick of LK (4 passes on 640x480 stereo frames) + " + OPT,
    6: "C6 stereo only, USE_IMU = 0 (euroc_stereo_config.yaml): per window one camera tick of LK (4 passes on 752x480 stereo frames) + " + OPT,
}
CAMERA = {1: None, 2: (752, 480, True), 3: (640, 480, False), 4: (640, 480, True), 5: (640, 480, True), 6: (752, 480, True)}


def config_dict(cid, distinct, copies, no_lk):
    """`config` of the JSON line -- the SAME dict from both arms (the driver compares them): it names the workload, not what a run measured."""
    cam = None if no_lk else CAMERA[cid]
    return {"workload": WORKLOADS[cid] + (" [--no-lk: window solve only]" if (no_lk and CAMERA[cid]) else ""), "config_id": cid,
            "batch_per_gpu": distinct * copies, "distinct_sequences": distinct, "perturbed_copies": copies, "lk_in_step": cam is not None,
            "camera": None if cam is None else {"image": [cam[0], cam[1]], "stereo": cam[2], "features": N_FEAT, "images_uploaded_per_tick_e2e": 2 if cam[2] else 1},
            "l2": "working set >> 126 MB L2 at this batch; no explicit flush"}


def make_windows(rank, distinct, copies, config_id=2):
    """Steady-state windows (window 1 of each sequence, prior from window 0 comes later) for this rank."""
    cfg = synth.make_config(config_id)
    seqs = [synth.Sequence(cfg, 1000 * rank + i, 12) for i in range(distinct)]
    first = [s.window(0) for s in seqs]
    return cfg, seqs, first


def replicate(seqs, priors, prev_states, copies, rank):
    probs, states = [], []
    rng = np.random.default_rng(977 + rank)
    for s, pr, prev in zip(seqs, priors, prev_states):
        p, st, _ = s.window(1, prior=pr, prev_state=prev)
        for c in range(copies):
            x = st.copy()
            if c:
                x[0:77] += rng.normal(0, 1e-3, 77)
                for i in range(11):
                    x[7 * i + 3: 7 * i + 7] /= np.linalg.norm(x[7 * i + 3: 7 * i + 7])
                x[abi.STATE_FIXED:] *= 1.0 + rng.normal(0, 0.02, len(x) - abi.STATE_FIXED)
            probs.append(p)
            states.append(x)
    return probs, states


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def rank_max(x, world, device="cuda"):
    """Slowest rank's figure (every rank gets it): the job is as fast as its slowest replica."""
    if world <= 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(world, per_rank_units, steps, seconds):
    """Whole-job units/s of `world` independent replicas (weak scaling: every rank processes per_rank_units per step)."""
    return world * per_rank_units * steps / seconds


def usable_cores():
    """Host threads this process can really run on: affinity mask, capped by the cgroup CPU quota (a container may see
    128 CPUs and be allowed 16)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            q, per = open(path).read().split()[:2]
            if q != "max":
                n = max(1, min(n, int(float(q) / float(per) + 0.5)))
        except Exception:
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            n = max(1, min(n, int(q / per + 0.5)))
    except Exception:
        pass
    return n


def bind_to_gpu_numa(index):
    """Pin this rank (its feeder lanes inherit the mask) to the CPUs of the NUMA node the GPU hangs off, BEFORE any page-locked slab is allocated, so that
    staging memory is first-touched on that node and H2D copies do not cross the socket interconnect.  Returns a small report for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("00000000:"):
            bdf = "0000:" + bdf[9:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return {"node": None, "note": "no NUMA affinity reported for the GPU"}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"node": node, "note": "node CPUs outside this process' affinity mask"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception as e:      # no sysfs / nvidia-smi: run unbound
        return {"node": None, "note": "unbound (%s)" % type(e).__name__}


def cpu_solve_many(vo, probs, states, seconds, threads):
    """Oracle (C port of the reference algorithm) on `threads` pthreads inside the C library for about `seconds`;
    each optimisation is single-threaded like Ceres' default.  Returns (solves/s, count, seconds)."""
    flags = [abi.MARGIN_OLD] * len(probs)
    n, dt = vo.optimization_throughput(probs, states, flags, threads, 1)         # calibration pass
    rate = n / dt
    repeat = max(1, int(round(seconds * rate / max(1, n))))
    n, dt = vo.optimization_throughput(probs, states, flags, threads, repeat)
    return n / dt, n, dt


# ------------------------------------------------------------------------------------------------ camera frames (LK)
IMG_W, IMG_H, N_FEAT = 752, 480, 150          # C2 / C6 camera; --config 3/4/5 switch to 640 x 480 (set_camera)


def set_camera(cid):
    global IMG_W, IMG_H
    cam = CAMERA[cid]
    if cam is not None:
        IMG_W, IMG_H = cam[0], cam[1]
    return cam


def make_scenes(rank, scenes):
    """Synthetic EuRoC-sized stereo scenes: band-limited noise texture seen in two consecutive ticks (sub-pixel image
    motion) by a left and a right camera (horizontal disparity).  Returns per scene the four images and the feature
    positions in both ticks (a jittered 15x10 grid: ~150 features per frame, BASELINE.json configs[1])."""
    from scipy import ndimage
    out = []
    for k in range(scenes):
        rng = np.random.default_rng(5000 + 131 * rank + k)
        tex = ndimage.gaussian_filter(rng.normal(size=(IMG_H + 64, IMG_W + 64)).astype(np.float32), 2.0)
        tex = (tex - tex.min()) / (tex.max() - tex.min()) * 255.0
        dx, dy, disp = rng.uniform(-4, 4), rng.uniform(-3, 3), rng.uniform(3, 12)

        def view(sx, sy):
            return np.ascontiguousarray(ndimage.shift(tex, (sy, sx), order=1, mode="reflect")[32:32 + IMG_H, 32:32 + IMG_W].astype(np.uint8))
        gx, gy = np.meshgrid(np.linspace(50, IMG_W - 50, 15), np.linspace(50, IMG_H - 50, 10))
        p0 = (np.stack([gx.ravel(), gy.ravel()], 1) + rng.uniform(-8, 8, (N_FEAT, 2))).astype(np.float32)
        out.append({"left": [view(0, 0), view(dx, dy)], "right": [view(-disp, 0), view(dx - disp, dy)],
                    "pts": [p0, (p0 + np.float32([dx, dy])).astype(np.float32)]})
    return out


class FrameFeed:
    """Host-side camera buffers of F streams (page-locked): tick t shows image t%2 of each stream's scene."""

    def __init__(self, ctx, scenes, streams, stereo=True):
        self.ctx, self.F, self.stereo = ctx, streams, stereo
        idx = np.arange(streams) % len(scenes)
        self.left = [np.ascontiguousarray(np.stack([scenes[i]["left"][t] for i in idx])) for t in (0, 1)]
        self.right = [np.ascontiguousarray(np.stack([scenes[i]["right"][t] for i in idx])) for t in (0, 1)] if stereo else []
        self.pts = [np.ascontiguousarray(np.stack([scenes[i]["pts"][t] for i in idx])) for t in (0, 1)]
        self.n = np.full(streams, N_FEAT, np.int32)
        if ctx is not None:
            for a in self.left + self.right:
                ctx.host_register(a)

    def tick_bytes(self):
        k = 2 if self.stereo else 1
        return self.left[0].nbytes * k + k * self.pts[0].nbytes + k * self.n.nbytes

    def tick_args(self, t, sl=slice(None), first=False):
        """keyword arguments of LkBatch.upload for camera tick t (image t % 2 of every stream; the previous tick's image is resident)"""
        kw = dict(cur=self.left[t][sl], prev_pts=self.pts[1 - t][sl], n_prev=self.n[sl])
        if first:
            kw["prev"] = self.left[1 - t][sl]
        if self.stereo:
            kw.update(right=self.right[t][sl], stereo_pts=self.pts[t][sl], n_stereo=self.n[sl])
        return kw

    def close(self):
        if self.ctx is not None:
            for a in self.left + self.right:
                self.ctx.host_unregister(a)


def cv_track_frame(cv2, prev, cur, right, p_prev, p_cur, stereo=True):
    """The four calcOpticalFlowPyrLK calls + status rules of one FeatureTracker::trackImage (feature_tracker.cpp:139-162, 240-251)."""
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    h, w = prev.shape
    c, st, _ = cv2.calcOpticalFlowPyrLK(prev, cur, p_prev.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    rp, rs, _ = cv2.calcOpticalFlowPyrLK(cur, prev, c, p_prev.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    c2, rp2 = c.reshape(-1, 2), rp.reshape(-1, 2)
    d = np.sqrt(((p_prev.astype(np.float64) - rp2) ** 2).sum(1))
    ci = np.rint(c2).astype(np.int64)
    inb = (ci[:, 0] >= 1) & (ci[:, 0] < w - 1) & (ci[:, 1] >= 1) & (ci[:, 1] < h - 1)
    st_t = (st.reshape(-1) > 0) & (rs.reshape(-1) > 0) & (d <= 0.5) & inb
    if not stereo:
        return c2, st_t, None, None
    r, st2, _ = cv2.calcOpticalFlowPyrLK(cur, right, p_cur.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    b, st3, _ = cv2.calcOpticalFlowPyrLK(right, cur, r, None, winSize=(21, 21), maxLevel=3)
    r2, b2 = r.reshape(-1, 2), b.reshape(-1, 2)
    d2 = np.sqrt(((p_cur.astype(np.float64) - b2) ** 2).sum(1))
    ri = np.rint(r2).astype(np.int64)
    inb2 = (ri[:, 0] >= 1) & (ri[:, 0] < w - 1) & (ri[:, 1] >= 1) & (ri[:, 1] < h - 1)
    st_s = (st2.reshape(-1) > 0) & (st3.reshape(-1) > 0) & (d2 <= 0.5) & inb2
    return c2, st_t, r2, st_s


def cpu_lk_many(scenes, seconds, threads, stereo=True):
    """OpenCV's LK (the reference's tracker) on `threads` host threads, one camera frame per task; returns (frames/s, frames, seconds)."""
    try:
        import cv2
    except Exception:
        return None, 0, 0.0
    from concurrent.futures import ThreadPoolExecutor
    cv2.setNumThreads(1 if threads > 1 else 1)

    def one(i):
        sc = scenes[i % len(scenes)]
        cv_track_frame(cv2, sc["left"][0], sc["left"][1], sc["right"][1], sc["pts"][0], sc["pts"][1], stereo)
    t0 = time.perf_counter()
    one(0)
    per = time.perf_counter() - t0
    n = max(threads, int(seconds / max(per, 1e-4)) * threads)
    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(n)))
    else:
        for i in range(n):
            one(i)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def load_peaks():
    """Roofline denominators: the driver's HBM copy bandwidth (MEASURED_PEAKS.json) and this repo's FP64 / integer-issue / shared-memory
    microbenchmarks measured on the same pool (profiles/r02a_microbench.json, source profiles/micro/microbench.cu)."""
    peaks = {"hbm_gbs": 6650.0, "hbm_source": "fallback 6650", "fp64_tflops": 37.1, "int_tops": 18.4, "smem_tbs": 37.0, "micro_source": "nominal"}
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"], peaks["hbm_source"] = float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, of measured)"
    except Exception:
        pass
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02a_microbench.json")))
        peaks.update({"fp64_tflops": max(float(d["fp64_fma_tflops"]), float(d["dmma_m8n8k4_tflops"])), "int_tops": float(d["imad_tops"]),
                      "smem_tbs": float(d["smem_ld_tbs"]), "micro_source": "profiles/r02a_microbench.json (measured on this pool's B200)"})
    except Exception:
        pass
    return peaks


def kernel_models(probs, pri, B, lk):
    """Algorithmic bytes (SURVEY 8(d) per-unit figures x units per launch) and FP64 FLOPs (FMA = 2; dense counts of the algebra
    each kernel owns) / integer sample operations of ONE launch, per kernel."""
    n_vis = sum(len(p.vis_type) for p in probs)
    n_lm = sum(p.num_landmarks for p in probs)
    n_imu = sum(len(p.imu_frame_i) for p in probs)
    n_wheel = sum(len(p.wheel_frame_i) for p in probs)
    pg = sum(p.state_size for p in probs)
    R = [sum(abi.block_tsize(b) for b in range(32) if (p.block_flags[b] & 1) and not (p.block_flags[b] & 2)) for p in probs]
    npri = [int(p.prior.n) if p.prior is not None else 0 for p in probs]
    nmax = max([int(q.n) for q in pri if q is not None and q.valid] + [0])
    sR2 = float(sum(r * r + r for r in R))
    m = {
        "lin_vis": (112.0 * n_vis + 8.0 * pg, 700.0 * n_vis),                       # ~350 FMA per projection factor with its Jacobians
        "lm_reduce": (0.0, 2.0 * 30 * n_vis),
        "lin_vis_lm": (112.0 * n_vis + 8.0 * pg, 760.0 * n_vis),
        "asm_items": (0.0, 2.0 * 2 * (21 + 21 + 36 + 12) * n_vis),
        "asm_pairs": (0.0, 2.0 * 2 * (21 + 21 + 36 + 12) * n_vis),
        "syrk": (0.0, 2.0 * 80 * 81 / 2 * n_lm),
        "syrk_mma": (0.0, 2.0 * 80 * 81 / 2 * n_lm),                                # useful FLOPs of the symmetric product (the tensor tiles compute 65 of 110 8x8 tiles)
        "lin_small": (2296.0 * n_imu + 624.0 * n_wheel + 8.0 * sum(n * n + 2 * n for n in npri), 2.0 * (15 * 15 * 31) * n_imu + 2.0 * 2 * sum(n * n for n in npri)),
        "solve": (8.0 * sR2 + 8.0 * pg, sum(r ** 3 / 3.0 + 6.0 * r * r for r in R) * 2.0 / 2.0 + 2.0 * 3 * 80 * n_lm + 2.0 * 15 * 465 * n_imu),
        "marg": (8.0 * B * (nmax * nmax + nmax + abi.STATE_FIXED), B * (4.0 / 3.0 + 3.0) * 2.0 * nmax ** 3),     # tridiagonalisation + eigenvectors + QL
        "marg_eig": (0.0, B * (4.0 / 3.0 + 3.0) * 2.0 * nmax ** 3),
        "marg_prep": (8.0 * sum(n * n + 2 * n for n in npri), 2.0 * B * ((nmax + 15) ** 2 * 15 + 80 * 80 * (nmax + 15) / 2)),   # Schur of the dropped 15 + landmark elimination through T0
        "marg_tri": (0.0, B * 2.0 * (4.0 / 3.0) * nmax ** 3 * 2.0),                 # Householder reduction + accumulation of the transformations
        "marg_ql": (0.0, B * 30.0 * 2.0 * nmax ** 2),                               # ~2 sweeps per eigenvalue on (d, e): a dependent chain, not a throughput kernel
        "marg_apply": (8.0 * B * (nmax * nmax + nmax), B * 6.0 * 0.8 * nmax ** 3 + B * 2.0 * nmax ** 3),      # ~0.8 n^2 logged rotations (6 flops each) on every one of the n rows + J = S V^T; calibrated on the FP64 pipe utilisation ncu reports (23 %)
        "pair_reduce": (0.0, 0.0),
        "pair_win": (0.0, 2.0 * 2 * (21 + 21 + 36 + 12) * n_vis),                    # asm_pairs + pair_reduce in one kernel: the same products
        "lk_track": ((lk.algorithmic_bytes() if lk is not None else 0.0), 0.0),
        "lk_pyr_down": (0.0, 0.0),
    }
    return m


# which resource binds each kernel, from its ncu capture (profiles/*.ncu.txt): "latency" = dependent chains / barriers at the occupancy the
# shared-memory or register footprint allows -- neither the byte nor the flop roof is near, both fractions are reported anyway
BOUND_HINT = {"lk_track": "issue", "lk_pyr_down": "hbm", "lin_vis": "hbm", "lin_vis_lm": "latency", "lin_vis_lm_wide": "latency", "lm_reduce": "hbm", "asm_items": "hbm",
              "asm_pairs": "fp64", "asm_pairs_wide": "fp64", "syrk": "fp64", "syrk_mma": "fp64", "solve": "latency", "pair_reduce": "latency", "pair_win": "fp64",
              "marg": "latency", "marg_eig": "latency", "marg_prep": "latency", "marg_tri": "latency", "marg_ql": "latency", "marg_apply": "fp64", "lin_small": "latency",
              "setup": "latency", "prior_setup": "fp64", "reanchor": "latency", "lk_post": "hbm"}


def roofline_report(prof, probs, pri, B, lk, alg_bytes, steps, ms):
    peaks = load_peaks()
    models = kernel_models(probs, pri, B, lk)
    tot = sum(v[0] for v in prof.values())
    traffic_db = {}
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    kernels = {}
    for k, v in prof.items():
        base = k[:-5] if k.endswith("_marg") else k
        ms_l = v[0] / max(1, v[1])
        ab, fl = models.get(base, models.get(base.replace("_wide", ""), (0.0, 0.0))) if not k.endswith("_marg") else (0.0, 0.0)
        gbs = ab / (ms_l * 1e-3) / 1e9 if ms_l > 0 else 0.0
        tfl = fl / (ms_l * 1e-3) / 1e12 if ms_l > 0 else 0.0
        ent = {"ms_per_launch": ms_l, "launches_per_step": v[1] / 2, "share": v[0] / tot, "hbm_gbs": gbs, "hbm_frac": gbs / peaks["hbm_gbs"],
               "fp64_tflops": tfl, "fp64_frac": tfl / peaks["fp64_tflops"]}
        ent["bound"] = BOUND_HINT.get(base, "latency")
        ent["nearest_roof"] = "fp64" if ent["fp64_frac"] > ent["hbm_frac"] else "hbm"     # the roof the two fractions put the kernel closest to
        tr = traffic_db.get(base)
        if tr and not k.endswith("_marg"):
            ent["dram_traffic_bytes"] = tr["bytes_per_unit"] * B
        kernels[k] = ent
    top = max(prof.items(), key=lambda kv: kv[1][0])[0]
    t = kernels[top]
    ab = models.get(top.replace("_marg", ""), (0.0, 0.0))[0]
    # the contract's two roofs: the byte roof (hbm) or the arithmetic one (here FP64: DFMA / DMMA, measured) -- whichever the kernel is closer to; `limiter`
    # says what actually binds it
    if t["nearest_roof"] == "fp64":
        roof = {"bound": "fp64", "achieved": t["fp64_tflops"], "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": t["fp64_frac"]}
    else:
        roof = {"bound": "hbm", "achieved": t["hbm_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": t["hbm_frac"]}
    roof.update({"kernel": top, "limiter": t["bound"], "traffic": t.get("dram_traffic_bytes"), "algorithmic_bytes_per_launch": ab, "hbm_gbs": t["hbm_gbs"], "hbm_frac": t["hbm_frac"],
                 "fp64_tflops": t["fp64_tflops"], "fp64_frac": t["fp64_frac"],
                 "peak_source": {"hbm": peaks["hbm_source"], "fp64_int_smem": peaks["micro_source"]},
                 "whole_step": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes * steps / (ms * 1e-3) / 1e9,
                                "frac_of_hbm": alg_bytes * steps / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}})
    return roof, kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="viwb")
    ap.add_argument("--distinct", type=int, default=37, help="distinct synthetic sequences per rank (37 x 48 copies = 1776 windows = 12 x 148 SMs: whole waves of every one-block-per-window kernel -- solve at 2 / SM, marg at 3 / SM, syrk at 4 / SM)")
    ap.add_argument("--copies", type=int, default=48, help="perturbed initial guesses per sequence (batch = distinct*copies)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--e2e-lanes", type=int, default=4, help="host threads (each with its own context) driving the e2e measurement")
    ap.add_argument("--e2e-lanes-sweep", default="", help="comma-separated host-thread counts to also measure e2e at (tuning aid)")
    ap.add_argument("--no-lk", action="store_true", help="window solve only (no feature-tracker work in the step)")
    ap.add_argument("--overlap-lk", action="store_true", help="run the camera tick on its own context + CUDA stream beside the solver (measured: no gain, r01za)")
    ap.add_argument("--scenes", type=int, default=8, help="distinct synthetic stereo scenes per rank (replicated over the streams)")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5, 6], help="BASELINE.json configuration C1..C5 (C6 = stereo only, USE_IMU = 0); the metric is quoted on C2 = the default")
    ap.add_argument("--parity-windows", type=int, default=0, help="windows compared with the oracle after the timed region (0 = the whole batch)")
    args = ap.parse_args()
    cid = args.config
    synth_cid = 4 if cid == 5 else cid          # C5 = C4-shaped sequences, one shard per GPU
    cam = set_camera(cid)
    if cam is None:
        args.no_lk = True
    stereo = bool(cam and cam[2])
    config = config_dict(cid, args.distinct, args.copies, args.no_lk)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        # the reference's own solver (Ceres) cannot be built here: the arm is the oracle port (its factor math is pinned to the reference's
        # compiled sources, oracle/_ref; the trust-region solve is restated), plus OpenCV's own LK
        if rank != 0:
            return
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import viw_oracle as vo
        cores = usable_cores()
        cfg, seqs, first = make_windows(0, min(args.distinct, 16), 1, synth_cid)
        priors, prev = [], []
        for (p, st, _) in first:
            a, sm, q = vo.optimization(p, st, abi.MARGIN_OLD)
            priors.append(q)
            prev.append(a)
        probs, states = replicate(seqs, priors, prev, max(1, (2 * cores) // max(1, len(seqs))), 0)
        per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + warmup)))
        for _ in range(warmup):
            vo.optimization_throughput(probs, states, [abi.MARGIN_OLD] * len(probs), cores, 1)
        tot_n, tot_t = 0, 0.0
        lk_n, lk_t = 0, 0.0
        scenes = None if args.no_lk else make_scenes(0, min(args.scenes, 4))
        for _ in range(args.steps):
            rate, n, dt = cpu_solve_many(vo, probs, states, per_step * (0.8 if scenes else 1.0), cores)
            tot_n += n
            tot_t += dt
            if scenes:
                lr, ln, ldt = cpu_lk_many(scenes, per_step * 0.2, cores, stereo)
                if lr is None:
                    scenes = None
                else:
                    lk_n += ln
                    lk_t += ldt
        solve_rate = tot_n / tot_t
        lk_rate = lk_n / lk_t if lk_n else None
        # one window = one camera tick of LK + one optimisation, both on the same host cores one after the other
        value = 1.0 / (1.0 / solve_rate + (1.0 / lk_rate if lk_rate else 0.0))
        sample = "%d oracle optimisations (C FP64 port of the reference algorithm, %d pthreads) over %.1f s = %.1f solves/s" % (tot_n, cores, tot_t, solve_rate)
        if lk_rate:
            sample += "; %d camera ticks of OpenCV calcOpticalFlowPyrLK x%d (%d threads) over %.1f s = %.1f frames/s; value = harmonic combination" % (lk_n, 4 if stereo else 2, cores, lk_t, lk_rate)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
                "ms_per_step": 1e3 * (tot_t + lk_t) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "sample": {"windows_per_step": tot_n // max(1, args.steps), "camera_ticks_per_step": lk_n // max(1, args.steps)},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from viwb import lib
    numa = bind_to_gpu_numa(local_rank) if world > 1 else {"node": None, "note": "single rank: unbound"}
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"           # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = lib.Context(local_rank)           # raises if the CUDA library / device is missing (no CPU path)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # ---- inputs: window 0 of every sequence is solved + marginalised on the GPU (untimed) to obtain the priors
    cfg, seqs, first = make_windows(rank, args.distinct, args.copies, synth_cid)
    p0 = [f[0] for f in first]
    s0 = [f[1] for f in first]
    a0, _, q0 = ctx.optimization_batch(p0, s0, [abi.MARGIN_OLD] * len(p0))
    probs, states = replicate(seqs, q0, a0, args.copies, rank)
    B = len(probs)
    flags = [abi.MARGIN_OLD] * B
    batch = ctx.batch(probs, states, flags)
    alg_bytes = batch.algorithmic_bytes()
    lk, feed, scenes, ctx_cam, cam_stream = None, None, None, None, None
    if not args.no_lk:
        scenes = make_scenes(rank, args.scenes)
        feed = FrameFeed(ctx, scenes, B, stereo)
        # The reference runs FeatureTracker::trackImage() in its own thread beside the estimator thread; --overlap-lk gives the camera tick
        # its own context + CUDA stream likewise.  Measured on B200 (profiles/r01za_bench_overlap.json vs r01za_bench_serial.json): 26 105 vs
        # 26 088 solves/s -- both kernel families already keep the SMs' issue slots busy, so the default stays one stream.
        ctx_cam = ctx
        if args.overlap_lk:
            ctx_cam = lib.Context(local_rank)
            cam_stream = torch.cuda.Stream()
            ctx_cam.set_stream(cam_stream.cuda_stream)
        lk = ctx_cam.lk_batch(B, IMG_W, IMG_H, N_FEAT, stereo=stereo, flow_back=True)
        # tick 0 (untimed): both left images resident, the resident "current" image is tick 1
        lk.upload(**feed.tick_args(1, first=True))
        lk.run()
        lk.download()
        alg_bytes += lk.algorithmic_bytes()

    def step():
        if lk is not None:
            lk.run()            # pyramids of the tick's new left/right images + 4 LK passes + status rules
        batch.run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    def all_launches():
        return ctx.launch_count() + (ctx_cam.launch_count() if ctx_cam is not None and ctx_cam is not ctx else 0)
    l0 = all_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if cam_stream is not None:
        cam_stream.wait_event(e0)           # the camera stream's first timed tick starts after e0 ...
    for _ in range(args.steps):
        step()
    if cam_stream is not None:
        cam_done = torch.cuda.Event()
        cam_done.record(cam_stream)
        stream.wait_event(cam_done)         # ... and e1 waits for its last one: the timed region covers K ticks + K solves
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = all_launches() - l0
    clocks = sampler.stop()
    ms = rank_max(ms, world)
    value = job_throughput(world, B, args.steps, ms * 1e-3)

    # ---- parity spot check of what was just timed (pose error vs the oracle on a few windows, rank 0)
    sts, sums, pri = batch.download()

    # ---- e2e: host buffers in / out through the C ABI.  `--e2e-lanes` host threads (default 2) each own a context, a slice of
    #      the sequences and their camera streams, and call the public entry points back to back; one lane's host-side lowering /
    #      copies overlap the other lane's kernels (contexts are independent; the C calls release the GIL).
    e2e_steps = max(1, min(args.steps, 5))
    lk_out = None
    if lk is not None:
        lk_out = [a.copy() for a in lk.download()]        # results of the timed configuration (tick 1), checked below
    def measure_e2e(want_lanes):
        lanes = max(1, min(want_lanes, B // 32 if B >= 64 else 1))
        bounds = [B * k // lanes for k in range(lanes + 1)]

        class Lane:
            def __init__(self, k):
                self.lo, self.hi = bounds[k], bounds[k + 1]
                self.ctx = ctx if k == 0 else lib.Context(local_rank)
                self.call = self.ctx.prepare_optimization_batch(probs[self.lo:self.hi], states[self.lo:self.hi], flags[self.lo:self.hi])
                self.lk = None
                if lk is not None:
                    n = self.hi - self.lo
                    self.lk = lk if lanes == 1 else self.ctx.lk_batch(n, IMG_W, IMG_H, N_FEAT, stereo=stereo, flow_back=True)
                    if lanes > 1:
                        self.lk.upload(**feed.tick_args(1, slice(self.lo, self.hi), first=True))
                        self.lk.run()
                        self.lk.download()
                    # the two alternating camera ticks with their arguments marshalled once (the same page-locked ring every tick)
                    self.up = [self.lk.prepare_upload(**feed.tick_args(t, slice(self.lo, self.hi))) for t in (0, 1)]
                self.tick = 0

            def step(self):
                if self.lk is not None:
                    self.up[self.tick % 2]()     # the camera delivers image t of every stream; the previous tick's image is resident
                    self.lk.run()         # asynchronous: overlaps with the host-side lowering of the windows below
                    self.tick += 1
                self.call()               # lowering + H2D + solve + re-anchor + marginalise + D2H (synchronises)
                if self.lk is not None:
                    self.lk.download()

            def close(self):
                if self.lk is not None and self.lk is not lk:
                    self.lk.close()
                if self.ctx is not ctx:
                    self.ctx.close()

        lane_objs = [Lane(k) for k in range(lanes)]

        def run_lanes(n):
            if lanes == 1:
                for _ in range(n):
                    lane_objs[0].step()
                return
            ths = [threading.Thread(target=lambda L=L: [L.step() for _ in range(n)]) for L in lane_objs]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        run_lanes(2)
        barrier()
        up0 = sum(L.ctx.h2d_bytes() for L in lane_objs)
        t0 = time.perf_counter()
        run_lanes(e2e_steps)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        measure_e2e.table_bytes = (sum(L.ctx.h2d_bytes() for L in lane_objs) - up0) / e2e_steps       # what the library put on the wire for the windows, per step
        for L in lane_objs:
            L.close()
        return e2e_s, lanes

    # the lowering inside each lane's call is itself multi-threaded: give every lane its share of the usable cores instead of 16 threads each
    os.environ.setdefault("VIWB_HOST_THREADS", str(max(2, usable_cores() // max(1, world) // max(1, args.e2e_lanes))))
    e2e_sweep = None
    if args.e2e_lanes_sweep:                # tuning aid: the same measurement at several host-thread counts (reported, not used for the headline)
        e2e_sweep = {}
        for n in [int(v) for v in args.e2e_lanes_sweep.split(",")]:
            sec, ln = measure_e2e(n)
            e2e_sweep[str(ln)] = job_throughput(world, B, e2e_steps, rank_max(sec, world))
    # host threads per rank: the default 4 lanes, but never more than this rank's share of the usable cores (8 ranks on a 16-core quota
    # would otherwise run 32 feeder threads), and at least 2 so that one lane's lowering still overlaps the other's kernels
    e2e_s, lanes = measure_e2e(max(2, min(args.e2e_lanes, usable_cores() // max(1, world))) if world > 1 else args.e2e_lanes)
    e2e_s = rank_max(e2e_s, world)
    e2e_value = job_throughput(world, B, e2e_steps, e2e_s)
    caller_tables = sum(p.vis_obs.nbytes + 4 * 4 * len(p.vis_type) + p.imu_data.nbytes + p.wheel_data.nbytes + 8 * p.state_size +
                        (8 * (p.prior.n ** 2 + p.prior.n + abi.STATE_FIXED) if p.prior is not None else 0) for p in probs)
    h2d = measure_e2e.table_bytes          # counted by the library from the slab it copies (viwb_h2d_bytes): the packed wire format of the window tables
    nmax = max(int(q.n) for q in pri if q is not None and q.valid) if any(q is not None and q.valid for q in pri) else 0
    d2h = sum(8 * p.state_size + 8 * (nmax * nmax + abi.MAX_PRIOR_DIM + abi.STATE_FIXED) + 4 * 67 for p in probs)
    if lk is not None:
        h2d += feed.tick_bytes()
        d2h += (2 if stereo else 1) * B * N_FEAT * (8 + 1)

    # ---- latency of ONE window through the host-buffer call (what a single robot sees per key-frame; not the headline metric)
    one = ctx.prepare_optimization_batch(probs[:1], states[:1], flags[:1])
    one(); one()
    t0 = time.perf_counter()
    for _ in range(10):
        one()
    single_ms = (time.perf_counter() - t0) * 100.0

    # ---- live per-kernel timing (CUDA events around every launch, separate pass so the headline is unperturbed)
    roofline, kernels = None, None
    if not args.no_profile:
        if cam_stream is not None:
            ctx_cam.set_stream(stream.cuda_stream)      # per-kernel times are taken with the launches serialised on one stream
        ctx.set_profiling(True)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        prof = ctx.profile()
        ctx.set_profiling(False)
        if cam_stream is not None:
            ctx_cam.set_stream(cam_stream.cuda_stream)
        roofline, kernels = roofline_report(prof, probs, pri, B, lk, alg_bytes, args.steps, ms)

    # ---- CPU baseline + parity on rank 0 (the oracle is only used here as checker / baseline, never on the product path)
    cpu_baseline, parity = None, None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import viw_oracle as vo
        cores = usable_cores()
        # every window of the timed batch against the oracle (all host cores): solved + re-anchored poses, iteration counts and the
        # order-independent content of the new prior (n, |J_lin|_F^2 = trace(J^T J), |J_lin^T r_lin|^2)
        npar = B if args.parity_windows <= 0 else min(B, args.parity_windows)
        pidx = list(range(B)) if npar == B else list(range(0, B, max(1, B // npar)))[:npar]
        o_states, o_iters, o_digest, o_dt = vo.optimization_many([probs[i] for i in pidx], [states[i] for i in pidx], [abi.MARGIN_OLD] * len(pidx), cores)
        ep_max = er_max = dig_max = 0.0
        it_same = n_same = 0
        for k, i in enumerate(pidx):
            ep, er = synth.pose_errors(o_states[k], sts[i])
            ep_max, er_max = max(ep_max, ep), max(er_max, er)
            it_same += int(o_iters[k] == sums[i].num_iterations)
            q = pri[i]
            if q is not None and q.valid:
                J, r = q.Jmat(), q.rvec()
                got = np.array([q.n, float((J * J).sum()), float(((J.T @ r) ** 2).sum())])
                n_same += int(got[0] == o_digest[k][0])
                dig_max = max(dig_max, float(np.max(np.abs(got[1:] - o_digest[k][1:]) / np.maximum(1e-300, np.abs(o_digest[k][1:])))))
            else:
                n_same += int(o_digest[k][0] == 0)
        parity = {"pose_err_m": ep_max, "pose_err_rad": er_max, "tolerance": [1e-4, 1e-4], "windows_checked": len(pidx), "of": B,
                  "iteration_counts_equal": it_same, "prior_dims_equal": n_same, "prior_information_rel_err_max": dig_max,
                  "oracle_pass": "%d windows on %d host threads in %.1f s" % (len(pidx), cores, o_dt)}
        lk_cv = None
        if lk is not None:
            try:
                import cv2
                from concurrent.futures import ThreadPoolExecutor
                cv2.setNumThreads(1)
                ref_by_scene = {}

                def ref_scene(k):
                    sc = scenes[k]
                    return cv_track_frame(cv2, sc["left"][0], sc["left"][1], sc["right"][1], sc["pts"][0], sc["pts"][1], stereo)
                with ThreadPoolExecutor(max(1, min(cores, len(scenes)))) as ex:
                    for k, res in enumerate(ex.map(ref_scene, range(len(scenes)))):
                        ref_by_scene[k] = res
                worst, agree, cnt = 0.0, 0.0, 0
                for f in range(B):          # every stream of the timed tick (streams replicate the scenes; the device results are per stream)
                    c, st_t, r, st_s = ref_by_scene[f % len(scenes)]
                    pairs = [(c, st_t, lk_out[0][f], lk_out[1][f])] + ([(r, st_s, lk_out[2][f], lk_out[3][f])] if stereo else [])
                    for ref_p, ref_s, got_p, got_s in pairs:
                        both = ref_s & (got_s > 0)
                        worst = max(worst, float(np.abs(ref_p[both] - got_p[both]).max()) if both.any() else 0.0)
                        agree += float((ref_s == (got_s > 0)).mean())
                        cnt += 1
                parity.update({"lk_max_px_vs_opencv": worst, "lk_status_agreement": agree / cnt, "lk_tolerance_px": 1e-2, "lk_streams_checked": B})
                lk_cv = cv2
            except ImportError:
                parity.update({"lk": "OpenCV not importable on this box: LK parity not checked in bench (see tests)"})
        if world == 1:
            rate, n, dt = cpu_solve_many(vo, probs, states, args.cpu_seconds * (0.8 if lk_cv else 1.0), 1)
            sample = "%d sequential oracle optimisations (C FP64 port of the reference algorithm; Ceres default num_threads=1) in %.1f s = %.1f solves/s" % (n, dt, rate)
            if lk_cv:
                lr, ln, ldt = cpu_lk_many(scenes, args.cpu_seconds * 0.2, 1, stereo)
                sample += "; %d camera ticks of OpenCV calcOpticalFlowPyrLK x%d (1 thread) in %.1f s = %.1f frames/s; value = harmonic combination" % (ln, 4 if stereo else 2, ldt, lr)
                rate = 1.0 / (1.0 / rate + 1.0 / lr)
            cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "workload_stats": {"mean_visual_factors": sum(len(p.vis_type) for p in probs) / B, "mean_landmarks": sum(p.num_landmarks for p in probs) / B,
                                   "camera_streams": B if lk is not None else 0, "distinct_scenes": len(scenes) if scenes else 0,
                                   "lk_stream": None if lk is None else ("same stream as the solver" if cam_stream is None else "own CUDA stream, concurrent with the solver (the reference's tracker thread)")},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "host_threads": lanes, "sweep": e2e_sweep, "caller_table_bytes_per_step": int(caller_tables),
                        "numa": numa, "h2d_gbs_per_rank": h2d * e2e_steps / e2e_s / 1e9,
                        "limiter": "host feed: %.2f GB of page-locked H2D per rank per step (camera frames + whole windows) against the device step" % (h2d / 1e9)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "parity": parity,
                "single_window_e2e_ms": single_ms}
        print(json.dumps(line))
    if lk is not None:
        lk.close()
        feed.close()
    batch.destroy()
    if ctx_cam is not None and ctx_cam is not ctx:
        ctx_cam.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

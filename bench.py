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


def cpu_solve_many(vo, probs, states, seconds, threads):
    """Oracle (C port of the reference algorithm) on `threads` pthreads inside the C library for about `seconds`;
    each optimisation is single-threaded like Ceres' default.  Returns (solves/s, count, seconds)."""
    flags = [abi.MARGIN_OLD] * len(probs)
    n, dt = vo.optimization_throughput(probs, states, flags, threads, 1)         # calibration pass
    rate = n / dt
    repeat = max(1, int(round(seconds * rate / max(1, n))))
    n, dt = vo.optimization_throughput(probs, states, flags, threads, repeat)
    return n / dt, n, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="viwb")
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic sequences per rank")
    ap.add_argument("--copies", type=int, default=32, help="perturbed initial guesses per sequence (batch = distinct*copies)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        # the reference's own CPU implementation cannot be built here (ROS/Ceres/Eigen absent): the arm is the oracle port
        if rank != 0:
            return
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import viw_oracle as vo
        cores = os.cpu_count() or 1
        cfg, seqs, first = make_windows(0, min(args.distinct, 16), 1)
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
        for _ in range(args.steps):
            rate, n, dt = cpu_solve_many(vo, probs, states, per_step, cores)
            tot_n += n
            tot_t += dt
        value = tot_n / tot_t
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
                "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": "C2 EuRoC-shaped stereo+IMU window with prior, 8 dogleg iterations + MARGIN_OLD marginalisation",
                                                "windows_per_sample": tot_n // max(1, args.steps)},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "%d oracle optimisations (C FP64 port, %d threads) over %.1f s" % (tot_n, cores, tot_t)},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from viwb import lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = lib.Context(local_rank)           # raises if the CUDA library / device is missing (no CPU path)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # ---- inputs: window 0 of every sequence is solved + marginalised on the GPU (untimed) to obtain the priors
    cfg, seqs, first = make_windows(rank, args.distinct, args.copies)
    p0 = [f[0] for f in first]
    s0 = [f[1] for f in first]
    a0, _, q0 = ctx.optimization_batch(p0, s0, [abi.MARGIN_OLD] * len(p0))
    probs, states = replicate(seqs, q0, a0, args.copies, rank)
    B = len(probs)
    flags = [abi.MARGIN_OLD] * B
    batch = ctx.batch(probs, states, flags)
    alg_bytes = batch.algorithmic_bytes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        batch.run()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        batch.run()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- parity spot check of what was just timed (pose error vs the oracle on a few windows, rank 0)
    sts, sums, pri = batch.download()

    # ---- e2e: host buffers in / out through the C ABI
    e2e_steps = max(1, min(args.steps, 5))
    e2e_call = ctx.prepare_optimization_batch(probs, states, flags)
    e2e_call()
    e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * e2e_steps / e2e_s
    h2d = sum(p.vis_obs.nbytes + 4 * 4 * len(p.vis_type) + p.imu_data.nbytes + p.wheel_data.nbytes + 8 * p.state_size +
              (8 * (p.prior.n ** 2 + p.prior.n + abi.STATE_FIXED) if p.prior is not None else 0) for p in probs)
    d2h = sum(8 * p.state_size + 8 * (abi.MAX_PRIOR_DIM ** 2 + abi.MAX_PRIOR_DIM + abi.STATE_FIXED) for p in probs)

    # ---- live per-kernel timing (CUDA events around every launch, separate pass so the headline is unperturbed)
    roofline, kernels = None, None
    if not args.no_profile:
        ctx.set_profiling(True)
        for _ in range(2):
            batch.run()
        torch.cuda.synchronize()
        prof = ctx.profile()
        ctx.set_profiling(False)
        tot = sum(v[0] for v in prof.values())
        kernels = {k: {"ms_per_launch": v[0] / max(1, v[1]), "launches_per_step": v[1] / 2, "share": v[0] / tot} for k, v in prof.items()}
        top = max(prof.items(), key=lambda kv: kv[1][0])[0]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic bytes of one launch of the dominant kernel: its share of the SURVEY 8(d) B_iter model (DESIGN.md "Roofline")
        n_vis = sum(len(p.vis_type) for p in probs)
        R = sum(sum(abi.block_tsize(b) for b in range(32) if (p.block_flags[b] & 1) and not (p.block_flags[b] & 2)) for p in probs) / B
        pg = sum(p.state_size for p in probs)
        per_launch = {"lin_vis": 112.0 * n_vis + 8.0 * pg, "lm_reduce": 0.0, "assemble": 8.0 * B * (R * R + R), "solve": 8.0 * B * (R * R + R) + 8.0 * pg,
                      "lin_small": sum(2296.0 * len(p.imu_frame_i) + 8.0 * (p.prior.n ** 2 + 2 * p.prior.n if p.prior is not None else 0) for p in probs)}
        a_bytes = per_launch.get(top.replace("_marg", ""), 0.0)
        achieved = a_bytes / (kernels[top]["ms_per_launch"] * 1e-3) / 1e9 if kernels[top]["ms_per_launch"] > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (sustained copy)" if peaks else "fallback 6650",
                    "algorithmic_bytes_per_launch": a_bytes,
                    "whole_step": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes * args.steps / (ms * 1e-3) / 1e9}}

    # ---- CPU baseline + parity on rank 0 (the oracle is only used here as checker / baseline, never on the product path)
    cpu_baseline, parity = None, None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import viw_oracle as vo
        ep_max = er_max = 0.0
        for i in range(0, B, max(1, B // 6)):
            a, sm, q = vo.optimization(probs[i], states[i], abi.MARGIN_OLD)
            ep, er = synth.pose_errors(a, sts[i])
            ep_max, er_max = max(ep_max, ep), max(er_max, er)
        parity = {"pose_err_m": ep_max, "pose_err_rad": er_max, "tolerance": [1e-4, 1e-4], "windows_checked": len(range(0, B, max(1, B // 6)))}
        if world == 1:
            rate, n, dt = cpu_solve_many(vo, probs, states, args.cpu_seconds, 1)
            cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": "%d sequential oracle optimisations (C FP64 port of the reference algorithm; Ceres default num_threads=1) in %.1f s" % (n, dt)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C2 EuRoC-shaped stereo+IMU window with prior (steady state): 8 dogleg iterations + gauge re-anchor + MARGIN_OLD marginalisation per solve",
                           "batch_per_gpu": B, "distinct_sequences": args.distinct, "perturbed_copies": args.copies,
                           "mean_visual_factors": sum(len(p.vis_type) for p in probs) / B, "mean_landmarks": sum(p.num_landmarks for p in probs) / B,
                           "l2": "working set >> 126 MB L2 at this batch; no explicit flush", "lk_in_step": False},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "parity": parity}
        print(json.dumps(line))
    batch.destroy()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

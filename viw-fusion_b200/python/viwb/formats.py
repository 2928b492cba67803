"""Reference-compatible on-disk formats: a synthetic Sequence written in the text layout the reference's
`bag_writer` tool turns into /sim/* topics, and solved poses written like `pubOdometry` writes vio.csv -- so that anyone with
ROS + Ceres can run the genuine reference on the very inputs this repository solves and compare trajectories.

Layouts (one record per line, whitespace separated):
  imu_pose.txt / imu_pose_noise.txt       t qw qx qy qz tx ty tz gx gy gz ax ay az      vins_estimator/src/tool/bagwriter.cpp:88-90
  wheel_pose.txt / wheel_pose_noise.txt   t qw qx qy qz tx ty tz gx gy gz vx vy vz      bagwriter.cpp:122-124
  imu_pose_tum_correspondence_cam.txt     t tx ty tz qx qy qz qw   (TUM)                 bagwriter.cpp:154-156
  keyframe/pixel[_noise]_all_points_<i>.txt   X Y Z 1 u v t   (one file per camera frame, u v in pixels)   bagwriter.cpp:176-201
  vio.csv                                 stamp_ns px py pz qx qy qz qw  (fixed, 0 / 9 decimals)           utility/visualization.cpp:295-307
`bag_writer` adds time_offset = 1.65e9 s to every stamp (bagwriter.cpp:34); the files hold sequence time.
The simulation front end (FeatureTracker::trackFeature, feature_tracker.cpp:332-372) assumes the same landmarks in the same order in
every frame, so only sequences whose landmarks are visible throughout (C1 shapes: SynthConfig.landmarks_all_frames) can be replayed through
it; ragged sequences are still written (one line per visible landmark) for tools that match by position.
"""
import os

import numpy as np

from .geom import R_to_q

TIME_OFFSET = 1.65e9      # bagwriter.cpp:34


def _q_wxyz(R):
    q = R_to_q(R)         # [x, y, z, w]
    return [q[3], q[0], q[1], q[2]]


def _write(path, rows):
    with open(path, "w") as f:
        for r in rows:
            f.write(" ".join(repr(float(v)) for v in r) + "\n")


def read_table(path):
    with open(path) as f:
        rows = [[float(v) for v in line.split()] for line in f if line.strip()]
    return np.array(rows)


def write_vio_data_simulation(seq, outdir):
    """Writes the seven inputs of bag_writer for a synth.Sequence. Returns {name: path}."""
    cfg, traj = seq.cfg, seq.traj
    os.makedirs(os.path.join(outdir, "keyframe"), exist_ok=True)
    out = {}
    # ---- IMU: the interval buffers share their end points; the file keeps every sample once (an interval's last sample is the next one's first)
    rows_n, rows_c = [], []
    for j, (dt, acc, gyr) in enumerate(seq.imu):
        ts = seq.t_frames[j] + np.r_[0.0, np.cumsum(dt)]
        last = j == len(seq.imu) - 1
        for k in range(len(ts) if last else len(ts) - 1):
            t = ts[k]
            R, p = traj.rot(t), traj.pos(t)
            head = [t] + _q_wxyz(R) + list(p)
            rows_n.append(head + list(gyr[k]) + list(acc[k]))
            rows_c.append(head + list(traj.omega_body(t)) + list(R.T @ (traj.acc(t) + seq.G)))
    out["imu_pose.txt"] = os.path.join(outdir, "imu_pose.txt"); _write(out["imu_pose.txt"], rows_c)
    out["imu_pose_noise.txt"] = os.path.join(outdir, "imu_pose_noise.txt"); _write(out["imu_pose_noise.txt"], rows_n)
    # ---- wheel odometer (body-frame velocity + angular rate in the odometer frame)
    if cfg.use_wheel:
        rows_n, rows_c = [], []
        for j, (dt, vel, gyr) in enumerate(seq.wheel):
            ts = seq.t_frames[j] + np.r_[0.0, np.cumsum(dt)]
            last = j == len(seq.wheel) - 1
            for k in range(len(ts) if last else len(ts) - 1):
                t = ts[k]
                Rwb, wb = traj.rot(t), traj.omega_body(t)
                Rwo, pwo = Rwb @ cfg.R_io, traj.pos(t) + Rwb @ cfg.t_io
                head = [t] + _q_wxyz(Rwo) + list(pwo)
                v_o = Rwo.T @ (traj.vel(t) + Rwb @ np.cross(wb, cfg.t_io))
                rows_n.append(head + list(gyr[k]) + list(vel[k]))
                rows_c.append(head + list(cfg.R_io.T @ wb) + list(v_o))
        out["wheel_pose.txt"] = os.path.join(outdir, "wheel_pose.txt"); _write(out["wheel_pose.txt"], rows_c)
        out["wheel_pose_noise.txt"] = os.path.join(outdir, "wheel_pose_noise.txt"); _write(out["wheel_pose_noise.txt"], rows_n)
    # ---- ground truth at the camera stamps, TUM order
    rows = []
    for f, t in enumerate(seq.t_frames):
        q = R_to_q(seq.gt_R[f])
        rows.append([t] + list(seq.gt_P[f]) + list(q))
    out["imu_pose_tum_correspondence_cam.txt"] = os.path.join(outdir, "imu_pose_tum_correspondence_cam.txt")
    _write(out["imu_pose_tum_correspondence_cam.txt"], rows)
    # ---- features per camera frame: world point (homogeneous), pixel, stamp
    for f, t in enumerate(seq.t_frames):
        clean, noisy = [], []
        for tr in seq.tracks:
            if not (tr["start"] <= f < tr["end"]):
                continue
            pl = tr["obs"][f - tr["start"]][1]
            Xc = seq._project(tr["X"], f, 0)
            head = list(tr["X"]) + [1.0]
            clean.append(head + [cfg.fx * Xc[0] / Xc[2] + cfg.cx, cfg.fy * Xc[1] / Xc[2] + cfg.cy, t])
            noisy.append(head + [cfg.fx * pl[0] + cfg.cx, cfg.fy * pl[1] + cfg.cy, t])
        _write(os.path.join(outdir, "keyframe", "pixel_all_points_%d.txt" % f), clean)
        _write(os.path.join(outdir, "keyframe", "pixel_noise_all_points_%d.txt" % f), noisy)
    out["keyframe"] = os.path.join(outdir, "keyframe")
    return out


def read_vio_data_simulation(outdir):
    """Reads what write_vio_data_simulation wrote (also the output of HeYijia's vio_data_simulation with the reference's pixel files)."""
    d = {}
    for name in ("imu_pose", "imu_pose_noise", "wheel_pose", "wheel_pose_noise"):
        p = os.path.join(outdir, name + ".txt")
        if os.path.exists(p):
            a = read_table(p)
            d[name] = {"t": a[:, 0], "q_wxyz": a[:, 1:5], "p": a[:, 5:8], "gyr": a[:, 8:11], "acc" if name.startswith("imu") else "vel": a[:, 11:14]}
    p = os.path.join(outdir, "imu_pose_tum_correspondence_cam.txt")
    if os.path.exists(p):
        a = read_table(p)
        d["groundtruth"] = {"t": a[:, 0], "p": a[:, 1:4], "q_xyzw": a[:, 4:8]}
    frames, i = [], 0
    while os.path.exists(os.path.join(outdir, "keyframe", "pixel_noise_all_points_%d.txt" % i)):
        a = read_table(os.path.join(outdir, "keyframe", "pixel_noise_all_points_%d.txt" % i))
        frames.append({"X": a[:, 0:3], "uv": a[:, 4:6], "t": float(a[-1, 6]) if len(a) else None})
        i += 1
    d["frames"] = frames
    return d


def format_vio_csv_line(stamp_s, p, q_xyzw):
    """One line of vio.csv exactly as pubOdometry prints it (visualization.cpp:295-307): ios::fixed, stamp*1e9 with 0 decimals, then
    position and quaternion (x y z w) with 9 decimals, space separated."""
    return "%.0f %s\n" % (stamp_s * 1e9, " ".join("%.9f" % float(v) for v in list(p) + list(q_xyzw)))


def write_vio_csv(path, stamps, states, frame=10, append=False):
    """states: solved window state vectors (include/viwb.h layout); writes the pose of `frame` (WINDOW_SIZE = the newest) per stamp."""
    with open(path, "a" if append else "w") as f:
        for t, st in zip(stamps, states):
            f.write(format_vio_csv_line(t, st[7 * frame: 7 * frame + 3], st[7 * frame + 3: 7 * frame + 7]))


def read_vio_csv(path):
    a = read_table(path)
    return {"t": a[:, 0] * 1e-9, "p": a[:, 1:4], "q_xyzw": a[:, 4:8]}


def ate_rmse(p_est, p_gt):
    """Absolute trajectory error after the optimal rigid alignment (Horn / Umeyama without scale), metres."""
    a, b = np.asarray(p_est, float), np.asarray(p_gt, float)
    ma, mb = a.mean(0), b.mean(0)
    U, _, Vt = np.linalg.svd((b - mb).T @ (a - ma))
    S = np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))])
    R = U @ S @ Vt
    return float(np.sqrt((((a - ma) @ R.T + mb - b) ** 2).sum(1).mean()))

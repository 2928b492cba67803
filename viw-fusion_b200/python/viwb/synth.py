"""Seeded synthetic generator for the BASELINE.json configurations.

Produces, for a configuration id C1..C4 and a sequence index, a short visual-inertial(-wheel) sequence
and lowers any 11-frame window of it to the `viwb_problem` tables of include/viwb.h:
  C1  mono + IMU, 100 landmarks seen in all 11 frames (1000 TwoFrameOneCam + 10 IMU factors)
  C2  EuRoC-shaped stereo + IMU, 150 tracked points / frame, geometric track lifetimes (mean 8 frames)
  C3  mono + IMU + wheel (D435i shapes), camera extrinsic (subset {2,6}) and td free
  C4  stereo + IMU + wheel + plane, everything shipped in the ridgeback config
  (C5 = several C4 sequences with different seeds, one per GPU)
  C6  stereo only, USE_IMU = 0 (config/euroc/euroc_stereo_config.yaml:5, config/kitti_odom/*.yaml:4): no speed-bias blocks,
      no IMU factors, para_Pose[0] constant (estimator.cpp:1398-1403)
RNG: numpy default_rng(1000*config + sequence).  The pre-integration below is this package's own
restatement of the upstream step that produces the factor constants (integration_base.h:63-167,
wheel_integration_base.h:67-177); it is not the oracle and does not import it.
"""
from dataclasses import dataclass, field

import numpy as np

from . import abi
from .geom import (skew, q_mul, q_normalize, q_to_R, R_to_q, so3_exp, so3_log, so3_Jr, Rz, Ry, Rx)


# ----------------------------------------------------------------------------- configuration
@dataclass
class SynthConfig:
    config_id: int = 1
    stereo: bool = False
    use_imu: bool = True
    use_wheel: bool = False
    use_plane: bool = False
    estimate_extrinsic: bool = False
    ex_subset_mask: int = 0            # bit i -> delta[i] frozen (PoseSubsetParameterization)
    estimate_td: bool = False
    estimate_ex_wheel: bool = False
    estimate_ix_wheel: bool = False
    width: int = 752
    height: int = 480
    fx: float = 461.16
    fy: float = 459.75
    cx: float = 376.0
    cy: float = 240.0
    acc_n: float = 0.1
    gyr_n: float = 0.01
    acc_w: float = 1e-3
    gyr_w: float = 1e-4
    g_norm: float = 9.81007
    vel_n_wheel: float = 0.01
    gyr_n_wheel: float = 0.004
    pitch_n: float = 0.01
    roll_n: float = 0.01
    zpw_n: float = 0.05
    planar: bool = False
    landmarks_all_frames: int = 0      # C1: this many landmarks visible in every frame
    tracks_per_frame: int = 150
    mean_track_len: float = 8.0
    cam_rate: float = 20.0
    imu_rate: float = 200.0
    wheel_rate: float = 50.0
    pixel_sigma: float = 1.0           # px, applied as N(0,(sigma/460)^2) in normalised coordinates
    outlier_frac: float = 0.02
    outlier_px: float = 10.0
    stereo_frac: float = 0.9
    R_ic: np.ndarray = field(default_factory=lambda: np.array([[0., 0., 1.], [-1., 0., 0.], [0., -1., 0.]]))
    t_ic0: np.ndarray = field(default_factory=lambda: np.array([0.041, 0.307, 0.544]))
    t_ic1: np.ndarray = field(default_factory=lambda: np.array([0.041, 0.258, 0.544]))
    R_io: np.ndarray = field(default_factory=lambda: np.eye(3))
    t_io: np.ndarray = field(default_factory=lambda: np.array([-0.208, 0.290, -0.168]))
    noise_scale: float = 0.3           # actual sensor noise = scale * nominal density (configs are inflated)


def make_config(cid):
    if cid == 1:   # config/euroc/euroc_mono_imu_config.yaml shapes
        return SynthConfig(config_id=1, landmarks_all_frames=100,
                           t_ic0=np.array([-0.0216, -0.0647, 0.0098]), t_ic1=np.array([-0.0198, 0.0454, 0.0079]))
    if cid == 2:   # config/euroc/euroc_stereo_imu_config.yaml shapes
        return SynthConfig(config_id=2, stereo=True,
                           t_ic0=np.array([-0.0216, -0.0647, 0.0098]), t_ic1=np.array([-0.0198, 0.0454, 0.0079]))
    if cid == 3:   # config/realsense_d435i/*: mono + IMU + wheel, ex (no z) and td free
        return SynthConfig(config_id=3, use_wheel=True, planar=True, estimate_extrinsic=True, ex_subset_mask=1 << 2,
                           estimate_td=True, width=640, height=480, fx=384.45, fy=384.45, cx=320.0, cy=240.0,
                           acc_n=0.1, gyr_n=0.05, acc_w=7.1765713730075628e-04, gyr_w=4.0e-05, g_norm=9.805)
    if cid in (4, 5):  # realsense_stereo_imu_config_ridgeback.yaml: stereo + IMU + wheel + plane
        return SynthConfig(config_id=cid, stereo=True, use_wheel=True, use_plane=True, planar=True,
                           estimate_extrinsic=True, ex_subset_mask=1 << 2,
                           width=640, height=480, fx=384.45, fy=384.45, cx=320.0, cy=240.0,
                           acc_n=0.1, gyr_n=0.05, acc_w=7.1765713730075628e-04, gyr_w=4.0e-05, g_norm=9.805)
    if cid == 6:   # config/euroc/euroc_stereo_config.yaml: imu 0, two cameras, extrinsics trusted
        return SynthConfig(config_id=6, stereo=True, use_imu=False,
                           t_ic0=np.array([-0.0216, -0.0647, 0.0098]), t_ic1=np.array([-0.0198, 0.0454, 0.0079]))
    raise ValueError("config id must be 1..6")


# ----------------------------------------------------------------------------- pre-integration
def imu_preintegrate(dt, acc, gyr, ba, bg, noise):
    """IntegrationBase::propagate over a sample buffer -> 287-double record (include/viwb.h)."""
    an, gn, aw, gw = noise
    Q = np.diag(np.r_[[an * an] * 3, [gn * gn] * 3, [an * an] * 3, [gn * gn] * 3, [aw * aw] * 3, [gw * gw] * 3])
    jac, cov = np.eye(15), np.zeros((15, 15))
    dp, dq, dv, sum_dt = np.zeros(3), np.array([0., 0., 0., 1.]), np.zeros(3), 0.0
    I3 = np.eye(3)
    for s in range(len(dt)):
        h, a0, g0, a1, g1 = dt[s], acc[s], gyr[s], acc[s + 1], gyr[s + 1]
        Rd = _qR(dq)
        un_gyr = 0.5 * (g0 + g1) - bg
        rq = q_mul(dq, np.r_[un_gyr * h / 2, 1.0])
        Rr = _qR(rq)
        un_acc_0 = _qrot(dq, a0 - ba)
        un_acc_1 = _qrot(rq, a1 - ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        rp = dp + dv * h + 0.5 * un_acc * h * h
        rv = dv + un_acc * h
        Rw, Ra0, Ra1 = skew(un_gyr), skew(a0 - ba), skew(a1 - ba)
        F = np.zeros((15, 15))
        F[0:3, 0:3] = I3
        F[0:3, 3:6] = -0.25 * Rd @ Ra0 * h * h + -0.25 * Rr @ Ra1 @ (I3 - Rw * h) * h * h
        F[0:3, 6:9] = I3 * h
        F[0:3, 9:12] = -0.25 * (Rd + Rr) * h * h
        F[0:3, 12:15] = -0.25 * Rr @ Ra1 * h * h * -h
        F[3:6, 3:6] = I3 - Rw * h
        F[3:6, 12:15] = -I3 * h
        F[6:9, 3:6] = -0.5 * Rd @ Ra0 * h + -0.5 * Rr @ Ra1 @ (I3 - Rw * h) * h
        F[6:9, 6:9] = I3
        F[6:9, 9:12] = -0.5 * (Rd + Rr) * h
        F[6:9, 12:15] = -0.5 * Rr @ Ra1 * h * -h
        F[9:12, 9:12] = I3
        F[12:15, 12:15] = I3
        V = np.zeros((15, 18))
        V[0:3, 0:3] = 0.25 * Rd * h * h
        V[0:3, 3:6] = 0.25 * -Rr @ Ra1 * h * h * 0.5 * h
        V[0:3, 6:9] = 0.25 * Rr * h * h
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * I3 * h
        V[3:6, 9:12] = 0.5 * I3 * h
        V[6:9, 0:3] = 0.5 * Rd * h
        V[6:9, 3:6] = 0.5 * -Rr @ Ra1 * h * 0.5 * h
        V[6:9, 6:9] = 0.5 * Rr * h
        V[6:9, 9:12] = V[6:9, 3:6]
        V[9:12, 12:15] = I3 * h
        V[12:15, 15:18] = I3 * h
        jac = F @ jac
        cov = F @ cov @ F.T + V @ Q @ V.T
        dp, dq, dv = rp, q_normalize(rq), rv
        sum_dt += h
    rec = np.zeros(abi.IMU_DOUBLES)
    rec[0], rec[1:4], rec[4:8], rec[8:11], rec[11:14], rec[14:17] = sum_dt, dp, dq, dv, ba, bg
    for k, (r0, c0) in enumerate(((0, 9), (0, 12), (3, 12), (6, 9), (6, 12))):
        rec[17 + 9 * k: 26 + 9 * k] = jac[r0:r0 + 3, c0:c0 + 3].reshape(-1)
    rec[62:] = cov.reshape(-1)
    return rec


def wheel_preintegrate(dt, vel, gyr, s, td, noise):
    """WheelIntegrationBase::propagate over a sample buffer -> 78-double record (include/viwb.h)."""
    vn, gn = noise
    Q = np.diag(np.r_[[vn * vn] * 3, [gn * gn] * 3, [vn * vn] * 3, [gn * gn] * 3])
    sx, sy, sw = s
    sv = np.diag([sx, sy, 1.0])
    jac, cov = np.zeros((6, 3)), np.zeros((6, 6))
    dp, dq, sum_dt = np.zeros(3), np.array([0., 0., 0., 1.]), 0.0
    for st in range(len(dt)):
        h, v0, g0, v1, g1 = dt[st], vel[st], gyr[st], vel[st + 1], gyr[st + 1]
        un_gyr = 0.5 * sw * (g0 + g1)
        ddq = np.r_[un_gyr * h / 2, 1.0]
        rq = q_mul(dq, ddq)
        Rd, Rr, Rdd = _qR(dq), _qR(rq), _qR(ddq)
        rp = dp + 0.5 * (_qrot(dq, sv @ v0) + _qrot(rq, sv @ v1)) * h
        Rv0, Rv1 = skew(sv @ v0), skew(sv @ v1)
        F = np.zeros((6, 6))
        F[0:3, 0:3] = np.eye(3)
        F[0:3, 3:6] = -0.5 * h * (Rd @ Rv0 + Rr @ Rv1 @ Rdd.T)
        F[3:6, 3:6] = Rdd.T
        Jr = so3_Jr(un_gyr * h)
        V = np.zeros((6, 12))
        V[0:3, 0:3] = 0.5 * h * Rd @ sv
        V[0:3, 3:6] = -0.25 * h * h * Rr @ Rv1 @ Jr
        V[0:3, 6:9] = 0.5 * h * Rr @ sv
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * Jr * sw * h
        V[3:6, 9:12] = 0.5 * Jr * sw * h
        I1, I2 = np.diag([1., 0, 0]), np.diag([0, 1., 0])
        jac[0:3, 0] += 0.5 * (Rd @ I1 @ v0 + Rr @ I1 @ v1) * h
        jac[0:3, 1] += 0.5 * (Rd @ I2 @ v0 + Rr @ I2 @ v1) * h
        last = jac[3:6, 2].copy()
        jac[3:6, 2] += Jr @ (0.5 * (g0 + g1)) * h
        jac[0:3, 2] += 0.5 * (Rd @ skew(last) @ sv @ v0 + Rr @ skew(jac[3:6, 2]) @ sv @ v1) * h
        cov = F @ cov @ F.T + V @ Q @ V.T
        dp, dq = rp, q_normalize(rq)
        sum_dt += h
    rec = np.zeros(abi.WHEEL_DOUBLES)
    rec[0:3], rec[3:7], rec[7:25], rec[25:61] = dp, dq, jac.reshape(-1), cov.reshape(-1)
    rec[61:64], rec[64] = [sx, sy, sw], td
    rec[65:68], rec[68:71], rec[71:74], rec[74:77], rec[77] = vel[0], gyr[0], vel[-1], gyr[-1], sum_dt
    return rec


def _qR(q):
    """Eigen toRotationMatrix formula applied to a possibly un-normalised quaternion (as the reference does)."""
    return q_to_R(q)


def _qrot(q, v):
    u = q[:3]
    uv = 2 * np.cross(u, v)
    return v + q[3] * uv + np.cross(u, uv)


# ----------------------------------------------------------------------------- trajectory
class Trajectory:
    def __init__(self, planar, rng):
        self.planar = planar
        self.ph = rng.uniform(0, 2 * np.pi, 6)
        self.t0 = rng.uniform(0.0, 4.0)

    def pos(self, t):
        t = t + self.t0
        if self.planar:   # lemniscate-like, z const
            return np.array([3.0 * np.sin(0.4 * t + self.ph[0]), 1.5 * np.sin(0.8 * t + 2 * self.ph[0]), 0.0])
        return np.array([1.5 * np.sin(0.8 * t + self.ph[0]), 1.0 * np.sin(1.1 * t + self.ph[1]), 0.3 * np.sin(1.7 * t + self.ph[2])])

    def vel(self, t):
        t = t + self.t0
        if self.planar:
            return np.array([1.2 * np.cos(0.4 * t + self.ph[0]), 1.2 * np.cos(0.8 * t + 2 * self.ph[0]), 0.0])
        return np.array([1.2 * np.cos(0.8 * t + self.ph[0]), 1.1 * np.cos(1.1 * t + self.ph[1]), 0.51 * np.cos(1.7 * t + self.ph[2])])

    def acc(self, t):
        t = t + self.t0
        if self.planar:
            return np.array([-0.48 * np.sin(0.4 * t + self.ph[0]), -0.96 * np.sin(0.8 * t + 2 * self.ph[0]), 0.0])
        return np.array([-0.96 * np.sin(0.8 * t + self.ph[0]), -1.21 * np.sin(1.1 * t + self.ph[1]), -0.867 * np.sin(1.7 * t + self.ph[2])])

    def rot(self, t):
        if self.planar:
            v = self.vel(t)
            return Rz(np.arctan2(v[1], v[0]))
        tt = t + self.t0
        return Rz(0.3 * np.sin(0.5 * tt + self.ph[3])) @ Ry(0.05 * np.sin(0.9 * tt + self.ph[4])) @ Rx(0.05 * np.sin(1.3 * tt + self.ph[5]))

    def omega_body(self, t, h=1e-5):
        return so3_log(self.rot(t - h).T @ self.rot(t + h)) / (2 * h)


# ----------------------------------------------------------------------------- sequence
class Sequence:
    """Ground truth + measurements of `num_frames` camera frames (>= 11)."""

    def __init__(self, cfg, seq=0, num_frames=11):
        self.cfg, self.seq, self.num_frames = cfg, seq, num_frames
        self.seed = 1000 * cfg.config_id + seq
        rng = np.random.default_rng(self.seed)
        self.rng = rng
        self.traj = Trajectory(cfg.planar, rng)
        self.dt_cam = 1.0 / cfg.cam_rate
        self.t_frames = np.arange(num_frames) * self.dt_cam
        self.G = np.array([0.0, 0.0, cfg.g_norm])
        self._make_states()
        self._make_imu()
        if cfg.use_wheel:
            self._make_wheel()
        self._make_tracks()

    # -- ground truth states
    def _make_states(self):
        cfg, rng = self.cfg, self.rng
        n = self.num_frames
        self.gt_P = np.array([self.traj.pos(t) for t in self.t_frames])
        self.gt_R = np.array([self.traj.rot(t) for t in self.t_frames])
        self.gt_V = np.array([self.traj.vel(t) for t in self.t_frames])
        ba0, bg0 = rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3)
        self.gt_Ba = ba0 + np.cumsum(rng.normal(0, cfg.acc_w * np.sqrt(self.dt_cam), (n, 3)), axis=0)
        self.gt_Bg = bg0 + np.cumsum(rng.normal(0, cfg.gyr_w * np.sqrt(self.dt_cam), (n, 3)), axis=0)

    def _make_imu(self):
        cfg, rng = self.cfg, self.rng
        per = int(round(cfg.imu_rate / cfg.cam_rate))
        h = self.dt_cam / per
        self.imu = []   # per interval j (frames j-1 -> j): (dt[per], acc[per+1,3], gyr[per+1,3])
        sa, sg = cfg.noise_scale * cfg.acc_n / np.sqrt(h), cfg.noise_scale * cfg.gyr_n / np.sqrt(h)
        for j in range(1, self.num_frames):
            ts = self.t_frames[j - 1] + np.arange(per + 1) * h
            acc = np.array([self.traj.rot(t).T @ (self.traj.acc(t) + self.G) for t in ts])
            gyr = np.array([self.traj.omega_body(t) for t in ts])
            w = (np.arange(per + 1) / per)[:, None]
            acc += (1 - w) * self.gt_Ba[j - 1] + w * self.gt_Ba[j] + rng.normal(0, sa, acc.shape)
            gyr += (1 - w) * self.gt_Bg[j - 1] + w * self.gt_Bg[j] + rng.normal(0, sg, gyr.shape)
            self.imu.append((np.full(per, h), acc, gyr))

    def _make_wheel(self):
        cfg, rng = self.cfg, self.rng
        hw = 1.0 / cfg.wheel_rate
        self.wheel = []
        for j in range(1, self.num_frames):
            t0, t1 = self.t_frames[j - 1], self.t_frames[j]
            ts = list(np.arange(t0, t1 - 1e-9, hw)) + [t1]
            ts = np.array(ts)
            vel, gyr = [], []
            for t in ts:
                Rwb, wb = self.traj.rot(t), self.traj.omega_body(t)
                v_o = (Rwb @ cfg.R_io).T @ (self.traj.vel(t) + Rwb @ np.cross(wb, cfg.t_io))
                vel.append(v_o)
                gyr.append(cfg.R_io.T @ wb)
            vel = np.array(vel) + rng.normal(0, cfg.noise_scale * cfg.vel_n_wheel, (len(ts), 3))
            gyr = np.array(gyr) + rng.normal(0, cfg.noise_scale * cfg.gyr_n_wheel, (len(ts), 3))
            self.wheel.append((np.diff(ts), vel, gyr))

    # -- landmarks and tracks
    def _cam_pose(self, f, cam):
        cfg = self.cfg
        t_ic = cfg.t_ic0 if cam == 0 else cfg.t_ic1
        return self.gt_R[f] @ cfg.R_ic, self.gt_P[f] + self.gt_R[f] @ t_ic

    def _project(self, X, f, cam):
        Rwc, pwc = self._cam_pose(f, cam)
        Xc = Rwc.T @ (X - pwc)
        return Xc

    def _in_fov(self, Xc):
        cfg = self.cfg
        if Xc[2] < 0.5:
            return False
        u, v = cfg.fx * Xc[0] / Xc[2] + cfg.cx, cfg.fy * Xc[1] / Xc[2] + cfg.cy
        return 5 <= u < cfg.width - 5 and 5 <= v < cfg.height - 5

    def _sample_landmark(self, f):
        cfg, rng = self.cfg, self.rng
        Rwc, pwc = self._cam_pose(f, 0)
        depth = rng.uniform(2.0, 15.0)
        u, v = rng.uniform(10, cfg.width - 10), rng.uniform(10, cfg.height - 10)
        Xc = depth * np.array([(u - cfg.cx) / cfg.fx, (v - cfg.cy) / cfg.fy, 1.0])
        return Rwc @ Xc + pwc

    def _make_tracks(self):
        """tracks: list of dict(X, start, obs=[(frame, ptL(3), ptR(3) or None)])."""
        cfg, rng, n = self.cfg, self.rng, self.num_frames
        tracks = []
        if cfg.landmarks_all_frames:
            while len(tracks) < cfg.landmarks_all_frames:
                X = self._sample_landmark(0)
                if all(self._in_fov(self._project(X, f, 0)) for f in range(n)):
                    tracks.append({"X": X, "start": 0, "end": n})
        else:
            active = []
            p_die = 1.0 / cfg.mean_track_len
            for f in range(n):
                alive = []
                for tr in active:
                    if rng.uniform() < p_die or not self._in_fov(self._project(tr["X"], f, 0)):
                        tr["end"] = f
                    else:
                        alive.append(tr)
                active = alive
                tries = 0
                while len(active) < cfg.tracks_per_frame and tries < 10000:
                    tries += 1
                    X = self._sample_landmark(f)
                    tr = {"X": X, "start": f, "end": n}
                    if f == 0:   # tracks older than the sequence start: stagger ages by killing some early
                        pass
                    active.append(tr)
                    tracks.append(tr)
        sig = cfg.pixel_sigma / 460.0
        for tr in tracks:
            obs = []
            for f in range(tr["start"], tr["end"]):
                XcL = self._project(tr["X"], f, 0)
                pl = np.array([XcL[0] / XcL[2], XcL[1] / XcL[2], 1.0])
                pl[:2] += rng.normal(0, sig, 2)
                if rng.uniform() < cfg.outlier_frac:
                    pl[:2] += rng.normal(0, cfg.outlier_px / 460.0, 2)
                pr = None
                if cfg.stereo and rng.uniform() < cfg.stereo_frac:
                    XcR = self._project(tr["X"], f, 1)
                    if self._in_fov(XcR):
                        pr = np.array([XcR[0] / XcR[2], XcR[1] / XcR[2], 1.0])
                        pr[:2] += rng.normal(0, sig, 2)
                obs.append((f, pl, pr))
            tr["obs"] = obs
        self.tracks = tracks

    # -- window lowering
    def gt_state(self, k, landmarks):
        """Ground-truth state vector of window k..k+10 for the given landmark list [(track, host_frame)]."""
        cfg = self.cfg
        st = np.zeros(abi.STATE_FIXED + len(landmarks))
        for i in range(abi.NUM_FRAMES):
            f = k + i
            st[7 * i: 7 * i + 3] = self.gt_P[f]
            st[7 * i + 3: 7 * i + 7] = R_to_q(self.gt_R[f])
            st[77 + 9 * i: 77 + 9 * i + 3] = self.gt_V[f]
            st[77 + 9 * i + 3: 77 + 9 * i + 6] = self.gt_Ba[f]
            st[77 + 9 * i + 6: 77 + 9 * i + 9] = self.gt_Bg[f]
        qic = R_to_q(cfg.R_ic)
        st[176:179], st[179:183] = cfg.t_ic0, qic
        st[183:186], st[186:190] = cfg.t_ic1, qic
        st[190:193], st[193:197] = cfg.t_io, R_to_q(cfg.R_io)
        st[197:201] = R_to_q(cfg.R_io)            # quirk 1: plane_R is filled from the wheel extrinsic quaternion
        st[201] = -(self.gt_P[k][2] + (self.gt_R[k] @ cfg.t_io)[2]) if cfg.planar else 0.0
        st[202:205] = 1.0
        st[205], st[206] = 0.0, 0.0
        for j, (tr, host) in enumerate(landmarks):
            Xc = self._project(tr["X"], host, 0)
            st[abi.STATE_FIXED + j] = 1.0 / Xc[2]
        return st

    def window(self, k=0, prior=None, prev_state=None, perturb=True, full=True):
        """Lower frames k..k+10 to (WindowProblem, initial_state, gt_state)."""
        cfg = self.cfg
        assert k + abi.NUM_FRAMES <= self.num_frames
        rng = np.random.default_rng(self.seed * 7919 + k)
        fc = abi.WINDOW_SIZE
        # landmarks with >= 4 observations inside the window (estimator.cpp:1589-1591)
        landmarks, vt, vl, vi, vj, vobs = [], [], [], [], [], []
        for tr in self.tracks:
            obs = [o for o in tr["obs"] if k <= o[0] <= k + fc]
            if len(obs) < 4:
                continue
            host = obs[0][0]
            idx = len(landmarks)
            landmarks.append((tr, host))
            pts_i = obs[0][1]
            vel_prev = {}
            prev = None
            vels, velsR = [], []
            for o in obs:   # tracker velocity: finite difference to the previous frame, 0 for a fresh feature
                if prev is not None and prev[0] == o[0] - 1:
                    vL = (o[1][:2] - prev[1][:2]) / self.dt_cam
                    vR = (o[2][:2] - prev[2][:2]) / self.dt_cam if (o[2] is not None and prev[2] is not None) else np.zeros(2)
                else:
                    vL, vR = np.zeros(2), np.zeros(2)
                vels.append(vL)
                velsR.append(vR)
                prev = o
            for o, vL, vR in zip(obs, vels, velsR):
                f = o[0]
                if f != host:
                    vt.append(abi.F_PROJ_2F1C); vl.append(idx); vi.append(host - k); vj.append(f - k)
                    vobs.append(np.r_[pts_i, o[1], vels[0], vL, 0.0, 0.0])
                if cfg.stereo and o[2] is not None:
                    vt.append(abi.F_PROJ_2F2C if f != host else abi.F_PROJ_1F2C)
                    vl.append(idx); vi.append(host - k); vj.append(f - k)
                    vobs.append(np.r_[pts_i, o[2], vels[0], vR, 0.0, 0.0])
        flags = np.zeros(abi.NUM_FIXED_BLOCKS, np.uint8)
        mask = np.zeros(abi.NUM_FIXED_BLOCKS, np.uint8)
        for i in range(abi.NUM_FRAMES):
            flags[abi.BLK_POSE0 + i] = abi.BLOCK_PRESENT
            if cfg.use_imu:
                flags[abi.BLK_SPEEDBIAS0 + i] = abi.BLOCK_PRESENT
        if not cfg.use_imu:
            flags[abi.BLK_POSE0] |= abi.BLOCK_CONSTANT        # estimator.cpp:1402-1403
        ncam = 2 if cfg.stereo else 1
        for c in range(ncam):
            flags[abi.BLK_EX_POSE0 + c] = abi.BLOCK_PRESENT | (0 if cfg.estimate_extrinsic else abi.BLOCK_CONSTANT)
            mask[abi.BLK_EX_POSE0 + c] = cfg.ex_subset_mask if cfg.estimate_extrinsic else 0
        if cfg.use_wheel:
            flags[abi.BLK_EX_WHEEL] = abi.BLOCK_PRESENT | (0 if cfg.estimate_ex_wheel else abi.BLOCK_CONSTANT)
            for b in (abi.BLK_SX, abi.BLK_SY, abi.BLK_SW):
                flags[b] = abi.BLOCK_PRESENT | (0 if cfg.estimate_ix_wheel else abi.BLOCK_CONSTANT)
        if cfg.use_plane:
            flags[abi.BLK_PLANE_R] = abi.BLOCK_PRESENT
            mask[abi.BLK_PLANE_R] = 1 << 2           # OrientationSubsetParameterization{2} (estimator.cpp:1497)
            flags[abi.BLK_PLANE_Z] = abi.BLOCK_PRESENT
        flags[abi.BLK_TD] = abi.BLOCK_PRESENT | (0 if cfg.estimate_td else abi.BLOCK_CONSTANT)
        flags[abi.BLK_TD_WHEEL] = abi.BLOCK_PRESENT | abi.BLOCK_CONSTANT
        imu_i, imu_j, imu_data = [], [], []
        noise = (cfg.acc_n, cfg.gyr_n, cfg.acc_w, cfg.gyr_w)
        gt = self.gt_state(k, landmarks)
        # initial state
        st = gt.copy()
        if perturb:
            for i in range(abi.NUM_FRAMES):
                st[7 * i: 7 * i + 3] += rng.normal(0, 0.05, 3)
                q = q_mul(st[7 * i + 3: 7 * i + 7], R_to_q(so3_exp(rng.normal(0, 0.01, 3))))
                st[7 * i + 3: 7 * i + 7] = q_normalize(q)
                st[77 + 9 * i: 77 + 9 * i + 3] += rng.normal(0, 0.05, 3)
                st[77 + 9 * i + 3: 77 + 9 * i + 6] += rng.normal(0, 0.005, 3)
                st[77 + 9 * i + 6: 77 + 9 * i + 9] += rng.normal(0, 0.0005, 3)
            st[abi.STATE_FIXED:] *= 1.0 + rng.normal(0, 0.1, len(landmarks))
            if cfg.estimate_extrinsic:
                for c in range(ncam):
                    o = 176 + 7 * c
                    st[o: o + 2] += rng.normal(0, 0.01, 2)
                    st[o + 3: o + 7] = q_normalize(q_mul(st[o + 3: o + 7], R_to_q(so3_exp(rng.normal(0, 0.005, 3)))))
        if prev_state is not None:   # carry the previous window's solution (frames k..k+9 were 1..10 there)
            st[0:7 * fc] = prev_state[7: 7 * (fc + 1)]
            st[77: 77 + 9 * fc] = prev_state[77 + 9: 77 + 9 * (fc + 1)]
            st[176:abi.STATE_FIXED] = prev_state[176:abi.STATE_FIXED]
        # factor constants are linearised at the *initial* bias estimates of the host frame (estimator.cpp:620-632)
        for j in range(1, abi.NUM_FRAMES if cfg.use_imu else 0):
            dt, acc, gyr = self.imu[k + j - 1]
            ba, bg = st[77 + 9 * j + 3: 77 + 9 * j + 6].copy(), st[77 + 9 * j + 6: 77 + 9 * j + 9].copy()
            imu_i.append(j - 1); imu_j.append(j)
            imu_data.append(imu_preintegrate(dt, acc, gyr, ba, bg, noise))
        wi, wj, wdata, pl = [], [], [], []
        if cfg.use_wheel:
            for j in range(1, abi.NUM_FRAMES):
                dt, vel, gyr = self.wheel[k + j - 1]
                wi.append(j - 1); wj.append(j)
                wdata.append(wheel_preintegrate(dt, vel, gyr, (1.0, 1.0, 1.0), 0.0, (cfg.vel_n_wheel, cfg.gyr_n_wheel)))
        if cfg.use_plane:
            pl = list(range(fc))
        g = abi.default_globals(cfg.g_norm, cfg.pitch_n, cfg.roll_n, cfg.zpw_n)
        prob = abi.WindowProblem(fc, len(landmarks), flags, mask, vt, vl, vi, vj, np.array(vobs) if vobs else np.zeros((0, 12)),
                                 imu_i, imu_j, np.array(imu_data) if imu_data else np.zeros((0, abi.IMU_DOUBLES)), wi, wj, np.array(wdata) if wdata else np.zeros((0, 78)),
                                 pl, prior, g)
        return prob, st, gt


def make_window(config_id=1, seq=0, k=0, num_frames=None, prior=None, perturb=True):
    cfg = make_config(config_id)
    s = Sequence(cfg, seq, num_frames if num_frames else k + abi.NUM_FRAMES)
    return s.window(k, prior=prior, perturb=perturb)


def pose_errors(state_a, state_b, frames=abi.NUM_FRAMES):
    """max position error [m] and max rotation error [rad] between two states over the window poses."""
    ep, er = 0.0, 0.0
    for i in range(frames):
        ep = max(ep, float(np.linalg.norm(state_a[7 * i: 7 * i + 3] - state_b[7 * i: 7 * i + 3])))
        Ra, Rb = q_to_R(q_normalize(state_a[7 * i + 3: 7 * i + 7])), q_to_R(q_normalize(state_b[7 * i + 3: 7 * i + 7]))
        er = max(er, float(np.linalg.norm(so3_log(Ra.T @ Rb))))
    return ep, er


def texture_image(h=480, w=752, seed=0):
    """Band-limited noise texture (two octaves), 8-bit: the synthetic camera image the detector / tracker tests look at."""
    from scipy import ndimage
    rng = np.random.default_rng(7000 + seed)
    t = ndimage.gaussian_filter(rng.normal(size=(h, w)).astype(np.float32), 2.0)
    t += 0.5 * ndimage.gaussian_filter(rng.normal(size=(h, w)).astype(np.float32), 6.0)
    return np.ascontiguousarray(((t - t.min()) / (t.max() - t.min()) * 255.0).astype(np.uint8))

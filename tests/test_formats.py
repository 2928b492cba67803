"""the reference-compatible text formats (viwb.formats) -- layout, round trip, and that the written IMU stream
means what bag_writer / the estimator take it to mean (body-frame specific force and rate: pre-integrating it reproduces the motion)."""
import os

import numpy as np

from viwb import abi, formats, synth


def test_write_read_round_trip(tmp_path):
    for cid in (1, 4):
        seq = synth.Sequence(synth.make_config(cid), 0, 12)
        out = formats.write_vio_data_simulation(seq, str(tmp_path / ("c%d" % cid)))
        d = formats.read_vio_data_simulation(str(tmp_path / ("c%d" % cid)))
        per = int(round(seq.cfg.imu_rate / seq.cfg.cam_rate))
        for name in ("imu_pose", "imu_pose_noise"):
            assert len(d[name]["t"]) == per * (seq.num_frames - 1) + 1 and np.all(np.diff(d[name]["t"]) > 0)
            assert np.allclose(np.linalg.norm(d[name]["q_wxyz"], axis=1), 1.0)
        # the noisy file holds the very samples the factors are pre-integrated from (interior samples of every interval)
        k = 3
        assert np.array_equal(d["imu_pose_noise"]["acc"][k * per: k * per + per], seq.imu[k][1][:per])
        assert np.array_equal(d["imu_pose_noise"]["gyr"][k * per: k * per + per], seq.imu[k][2][:per])
        assert ("wheel_pose.txt" in out) == seq.cfg.use_wheel
        if seq.cfg.use_wheel:
            assert np.array_equal(d["wheel_pose_noise"]["vel"][: len(seq.wheel[0][1]) - 1], seq.wheel[0][1][:-1])
        g = d["groundtruth"]
        assert np.array_equal(g["p"], seq.gt_P) and np.allclose(g["t"], seq.t_frames)
        assert len(d["frames"]) == seq.num_frames
        for f, fr in enumerate(d["frames"]):
            vis = [tr for tr in seq.tracks if tr["start"] <= f < tr["end"]]
            assert len(fr["uv"]) == len(vis) and abs(fr["t"] - seq.t_frames[f]) < 1e-12
            assert np.array_equal(fr["X"], np.array([tr["X"] for tr in vis]))
        if cid == 1:        # trackFeature's contract: the same landmarks, in the same order, in every frame
            assert all(np.array_equal(fr["X"], d["frames"][0]["X"]) for fr in d["frames"])


def test_clean_imu_stream_reproduces_the_motion(tmp_path):
    seq = synth.Sequence(synth.make_config(1), 3, 12)
    formats.write_vio_data_simulation(seq, str(tmp_path))
    imu = formats.read_vio_data_simulation(str(tmp_path))["imu_pose"]
    per = int(round(seq.cfg.imu_rate / seq.cfg.cam_rate))
    for j in (0, 5):
        sl = slice(j * per, (j + 1) * per + 1)
        rec = synth.imu_preintegrate(np.diff(imu["t"][sl]), imu["acc"][sl], imu["gyr"][sl], np.zeros(3), np.zeros(3), (0.1, 0.01, 1e-3, 1e-4))
        dt = rec[0]
        R0 = seq.gt_R[j]
        alpha = R0.T @ (seq.gt_P[j + 1] - seq.gt_P[j] - seq.gt_V[j] * dt + 0.5 * seq.G * dt * dt)
        beta = R0.T @ (seq.gt_V[j + 1] - seq.gt_V[j] + seq.G * dt)
        assert np.abs(rec[1:4] - alpha).max() < 1e-5 and np.abs(rec[8:11] - beta).max() < 1e-4


def test_vio_csv_matches_pubOdometry_formatting(tmp_path):
    assert formats.format_vio_csv_line(1.5, [1, 2, 3], [0, 0, 0, 1]) == \
        "1500000000 1.000000000 2.000000000 3.000000000 0.000000000 0.000000000 0.000000000 1.000000000\n"
    assert formats.format_vio_csv_line(1403636580.838555574, [-0.1234567894, 0, 0], [0, 0, 0, 1]).split()[:2] == ["1403636580838555648", "-0.123456789"]
    seq = synth.Sequence(synth.make_config(2), 0, 14)
    states, stamps = [], []
    for k in range(3):
        prob, st, gt = seq.window(k)
        states.append(gt); stamps.append(seq.t_frames[k + 10] + formats.TIME_OFFSET)
    p = str(tmp_path / "vio.csv")
    formats.write_vio_csv(p, stamps, states)
    r = formats.read_vio_csv(p)
    assert np.abs(r["p"] - seq.gt_P[10:13]).max() < 1e-9 and np.abs(r["t"] - np.array(stamps)).max() < 1e-6
    assert formats.ate_rmse(r["p"], seq.gt_P[10:13]) < 1e-9
    assert formats.ate_rmse(r["p"] + [1.0, -2.0, 0.5], seq.gt_P[10:13]) < 1e-9          # alignment removes a rigid offset

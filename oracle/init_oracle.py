"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the initialisation's visual-inertial(-wheel) alignment, SURVEY 8 f-4 ii:
solveGyroscopeBias (initial/initial_aligment.cpp:14-48), TangentBasis (:51-64), LinearAlignment / RefineGravity (:66-203),
LinearAlignmentWithWheel / RefineGravityWithWheel (:204-334).  Pinned against the reference's own compiled initial_aligment.cpp
(oracle/_ref, tests/test_reference_factors.py::test_visual_imu_alignment_restatement_matches_reference_code); `A.ldlt().solve(b)` is
numpy's LU solve here (Eigen is not in the image; the systems are well conditioned, agreement ~1e-10 relative)."""
import numpy as np

IMU, WHEEL = 287, 78


def quat_from_R(M):
    """Eigen::Quaterniond(Matrix3d) (x, y, z, w)"""
    t = M[0, 0] + M[1, 1] + M[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        return np.array([(M[2, 1] - M[1, 2]) * t, (M[0, 2] - M[2, 0]) * t, (M[1, 0] - M[0, 1]) * t, w])
    i = 0
    if M[1, 1] > M[0, 0]:
        i = 1
    if M[2, 2] > M[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    t = np.sqrt(M[i, i] - M[j, j] - M[k, k] + 1.0)
    q = np.zeros(4)
    q[i] = 0.5 * t
    t = 0.5 / t
    q[3] = (M[k, j] - M[j, k]) * t
    q[j] = (M[j, i] + M[i, j]) * t
    q[k] = (M[k, i] + M[i, k]) * t
    return q


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def solve_gyroscope_bias(R, imu):
    """:14-37 -> delta_bg.  R [F,3,3]; imu [F-1,287] (record i = pre_integration of frame i+1)"""
    A, b = np.zeros((3, 3)), np.zeros(3)
    for i in range(len(R) - 1):
        q_ij = quat_from_R(R[i].T @ R[i + 1])
        J = imu[i][35:44].reshape(3, 3)
        dq = imu[i][4:8]
        inv = np.array([-dq[0], -dq[1], -dq[2], dq[3]]) / (dq @ dq)
        tb = 2.0 * qmul(inv, q_ij)[:3]
        A += J.T @ J
        b += J.T @ tb
    return np.linalg.solve(A, b)


def tangent_basis(g0):
    a = g0 / np.linalg.norm(g0)
    tmp = np.array([0.0, 0.0, 1.0])
    if np.array_equal(a, tmp):
        tmp = np.array([1.0, 0.0, 0.0])
    b = tmp - a * (a @ tmp)
    b = b / np.linalg.norm(b)
    return np.stack([b, np.cross(a, b)], axis=1)


def _rows(i, R, T, imu, wheel, tic, rio, tio, lxly, g0):
    Ri, Rj = R[i], R[i + 1]
    dt, dp, dv = imu[i][0], imu[i][1:4], imu[i][8:11]
    rows = 9 if wheel is not None else 6
    cols = 10 if lxly is None else 9
    tA, tb = np.zeros((rows, cols)), np.zeros(rows)
    tA[0:3, 0:3] = -dt * np.eye(3)
    tA[0:3, cols - 1] = Ri.T @ (T[i + 1] - T[i]) / 100.0
    tb[0:3] = dp + Ri.T @ Rj @ tic - tic
    tA[3:6, 0:3] = -np.eye(3)
    tA[3:6, 3:6] = Ri.T @ Rj
    tb[3:6] = dv
    if lxly is None:
        tA[0:3, 6:9] = Ri.T * (dt * dt / 2)
        tA[3:6, 6:9] = Ri.T * dt
    else:
        tA[0:3, 6:8] = Ri.T * (dt * dt / 2) @ lxly
        tA[3:6, 6:8] = Ri.T * dt @ lxly
        tb[0:3] -= Ri.T * (dt * dt / 2) @ g0
        tb[3:6] -= Ri.T * dt @ g0
    if wheel is not None:
        tA[6:9, cols - 1] = (Ri @ rio).T @ (T[i + 1] - T[i]) / 100
        tb[6:9] = wheel[i][0:3] - rio.T @ Ri.T @ Rj @ tio + (Ri @ rio).T @ Rj @ tic - rio.T @ (tic - tio)
    return tA, tb


def _accumulate(A, b, i, tA, tb):
    n, cols = len(b), tA.shape[1]
    tail = cols - 6
    rA, rb = tA.T @ tA, tA.T @ tb
    A[3 * i:3 * i + 6, 3 * i:3 * i + 6] += rA[:6, :6]
    b[3 * i:3 * i + 6] += rb[:6]
    A[n - tail:, n - tail:] += rA[6:, 6:]
    b[n - tail:] += rb[6:]
    A[3 * i:3 * i + 6, n - tail:] += rA[:6, 6:]
    A[n - tail:, 3 * i:3 * i + 6] += rA[6:, :6]


def linear_alignment(R, T, imu, wheel, tic, rio, tio, g_norm):
    """LinearAlignment[WithWheel] + RefineGravity[WithWheel] -> (aligned, g, x); wheel None = the camera + IMU form"""
    F = len(R)
    n = 3 * F + 4
    A, b = np.zeros((n, n)), np.zeros(n)
    for i in range(F - 1):
        _accumulate(A, b, i, *_rows(i, R, T, imu, wheel, tic, rio, tio, None, None))
    A *= 1000.0
    b *= 1000.0
    x = np.linalg.solve(A, b)
    s = x[n - 1] / 100.0
    g = x[n - 4:n - 1].copy()
    if abs(np.linalg.norm(g) - g_norm) > 0.5 or s < 0:
        return False, g, x
    g0 = g / np.linalg.norm(g) * g_norm
    n = 3 * F + 3
    A, b = np.zeros((n, n)), np.zeros(n)                  # cleared once: the four passes below accumulate onto 1000x the previous system (:138-140)
    for _ in range(4):
        lxly = tangent_basis(g0)
        for i in range(F - 1):
            _accumulate(A, b, i, *_rows(i, R, T, imu, wheel, tic, rio, tio, lxly, g0))
        A *= 1000.0
        b *= 1000.0
        x = np.linalg.solve(A, b)
        g0 = g0 + lxly @ x[n - 3:n - 1]
        g0 = g0 / np.linalg.norm(g0) * g_norm
    s = x[n - 1] / 100.0
    x = x.copy()
    x[n - 1] = s
    return s >= 0.0, g0, x

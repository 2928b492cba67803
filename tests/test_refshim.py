"""The header stand-ins of oracle/refshim/ on their own: mini_eigen.h (inverse, LLT, products with run-time sized blocks, symmetric eigen
solver, quaternion algebra) and mini_sophus.h (SO3 exp / log) against numpy / scipy -- so that the agreement between the compiled
reference sources and the oracle (tests/test_reference_factors.py) cannot stem from a mistake the two share through the matrix header."""
import ctypes as C

import numpy as np
import pytest

vr = pytest.importorskip("viw_ref")
if not vr.available():
    pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libviw_ref.so", allow_module_level=True)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def test_inverse_llt_and_block_products():
    rng = np.random.default_rng(41)
    for trial in range(5):
        M = rng.normal(size=(15, 15))
        A = M @ M.T + 15 * np.eye(15) * 10.0 ** (-trial)               # SPD, condition number growing with the trial
        inv, llt, prod = np.zeros((15, 15)), np.zeros((15, 15)), np.zeros((15, 15))
        vr.lib().ref_selftest_linalg15(dp(np.ascontiguousarray(A)), dp(inv), dp(llt), dp(prod))
        assert np.abs(inv @ A - np.eye(15)).max() < 1e-9 * np.linalg.cond(A)
        assert np.abs(llt - np.linalg.cholesky(A)).max() < 1e-12 * np.abs(A).max() and np.allclose(np.triu(llt, 1), 0)
        F = np.zeros((15, 15)); F[3:6, 6:9] = A[0:3, 0:3]; F[0:3, 0:3] = 2 * np.eye(3)
        assert np.abs(prod - (F @ A @ F.T + A.T)).max() < 1e-12 * np.abs(prod).max()


@pytest.mark.parametrize("n", [3, 16, 76])
def test_symmetric_eigen_solver(n):
    rng = np.random.default_rng(42 + n)
    M = rng.normal(size=(n, n // 2 + 1))
    A = M @ M.T                                                      # rank deficient on purpose: the marginalization's systems are
    w, V = np.zeros(n), np.zeros((n, n))
    vr.lib().ref_selftest_eig(C.c_int(n), dp(np.ascontiguousarray(A)), dp(w), dp(V))
    w0 = np.linalg.eigvalsh(A)
    assert np.all(np.diff(w) >= -1e-12) and np.abs(w - w0).max() < 1e-12 * np.abs(w0).max()
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12 and np.abs(V @ np.diag(w) @ V.T - A).max() < 1e-11 * np.abs(A).max()


def test_quaternions_and_so3():
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(43)
    for trial in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        p = rng.normal(size=4); p /= np.linalg.norm(p)
        v = rng.normal(size=3) * (1e-12 if trial < 3 else 10.0 ** rng.uniform(-6, 0.4))
        out = np.zeros(31)
        vr.lib().ref_selftest_rotations(dp(q), dp(p), dp(v), dp(out))
        R, qi, qv, qr, qp, ex, lg = out[0:9].reshape(3, 3), out[9:13], out[13:16], out[16:20], out[20:24], out[24:28], out[28:31]
        R0 = Rot.from_quat(q).as_matrix()
        assert np.abs(R - R0).max() < 1e-14
        assert np.abs(qi - np.r_[-q[:3], q[3]]).max() < 1e-15
        assert np.abs(qv - R0 @ v).max() < 1e-14 * max(1.0, np.abs(v).max())
        assert min(np.abs(qr - q).max(), np.abs(qr + q).max()) < 1e-14
        qp0 = (Rot.from_quat(q) * Rot.from_quat(p)).as_quat()
        assert min(np.abs(qp - qp0).max(), np.abs(qp + qp0).max()) < 1e-14
        ex0 = Rot.from_rotvec(v).as_quat()
        assert min(np.abs(ex - ex0).max(), np.abs(ex + ex0).max()) < 1e-14
        lg0 = Rot.from_quat(q).as_rotvec()
        # SO3::log does not wrap to the shorter rotation: q and -q give rotation vectors 2 pi apart along the same axis
        alt = lg0 - 2 * np.pi * lg0 / max(np.linalg.norm(lg0), 1e-300)
        assert min(np.abs(lg - lg0).max(), np.abs(lg - alt).max()) < 1e-12

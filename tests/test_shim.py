"""The ceres:: shim (viw-fusion_b200/host): a translation unit shaped like Estimator::optimization() compiles against it,
lowers to the C ABI and reproduces what a direct C-ABI call gives.  `not gpu`: linked against the kernel-logic emulation;
`gpu`: against libviwb.so."""
import os
import struct
import subprocess

import numpy as np
import pytest

import parity_checks as pc
from viwb import abi, synth, lib as viwb_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "shim", "optimization_shim_test.cpp")


def build_exe(libpath, tag):
    exe = os.path.join(ROOT, "tests", "shim", "optimization_shim_test_" + tag)
    libdir, libname = os.path.dirname(libpath), os.path.basename(libpath)[3:-3]
    deps = [SRC, libpath] + [os.path.join(ROOT, "viw-fusion_b200", "host", d, f) for d in ("ceres", "factor", ".") for f in os.listdir(os.path.join(ROOT, "viw-fusion_b200", "host", d)) if f.endswith(".h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "viw-fusion_b200", "host"), SRC, "-o", exe, "-L", libdir, "-l" + libname,
                               "-Wl,-rpath," + libdir, "-pthread"])
    return exe


def wr(f, arr, fmt):
    a = np.ascontiguousarray(arr, fmt)
    f.write(struct.pack("i", a.size))
    f.write(a.tobytes())


def dump(path, cfg, prob, st, margin_flag):
    with open(path, "wb") as f:
        pr = prob.prior
        hdr = [prob.frame_count, prob.num_landmarks, 2 if cfg.stereo else 1, int(cfg.use_wheel), int(cfg.use_plane), int(cfg.estimate_extrinsic),
               int(cfg.ex_subset_mask), int(cfg.estimate_td), int(pr is not None), margin_flag]
        wr(f, hdr, np.int32)
        wr(f, st, np.float64); wr(f, prob.vis_obs, np.float64); wr(f, prob.imu_data, np.float64); wr(f, prob.wheel_data, np.float64)
        g = prob.globals
        wr(f, list(g.G) + list(g.vis_sqrt_info) + list(g.plane_sqrt_info), np.float64)
        wr(f, prob.vis_type, np.int32); wr(f, prob.vis_landmark, np.int32); wr(f, prob.vis_frame_i, np.int32); wr(f, prob.vis_frame_j, np.int32)
        wr(f, prob.block_flags, np.int32)
        if pr is not None:
            wr(f, [pr.n], np.int32); wr(f, [b for b, _ in pr.blocks()], np.int32); wr(f, [i for _, i in pr.blocks()], np.int32)
            wr(f, pr.x0, np.float64); wr(f, pr.Jmat(), np.float64); wr(f, pr.rvec(), np.float64)
        else:
            for fmt in (np.int32, np.int32, np.int32, np.float64, np.float64, np.float64):
                wr(f, [], fmt)


def rd(f):
    n = struct.unpack("i", f.read(4))[0]
    return np.frombuffer(f.read(8 * n), np.float64).copy()


def run_case(ctx, exe, tmp_path, cid, with_prior):
    cfg = synth.make_config(cid)
    seq = synth.Sequence(cfg, 0, 12)
    prob, st, _ = seq.window(0)
    if with_prior:
        a, _, q = ctx.optimization(prob, st, abi.MARGIN_OLD)
        prob, st, _ = seq.window(1, prior=q, prev_state=a)
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    dump(inp, cfg, prob, st, abi.MARGIN_OLD)
    subprocess.check_call([exe, inp, outp])
    with open(outp, "rb") as f:
        x = rd(f)
        meta = rd(f)
        ids = rd(f).reshape(-1, 3)
        J = rd(f)
        r = rd(f)
    # direct C-ABI reference: solve, then marginalise the *solved* state (the shim test does not re-anchor, like ceres::Solve)
    x_ref, sm = ctx.window_solve(prob, st)
    assert int(meta[0]) == sm.num_iterations
    ep, er = synth.pose_errors(x, x_ref)
    assert ep < 1e-9 and er < 1e-9
    assert np.abs(x - x_ref).max() < 1e-9
    q_ref = ctx.marginalize(prob, x_ref, abi.MARGIN_OLD)
    n = int(meta[2])
    assert meta[3] == 1.0 and n == q_ref.n
    assert [int(b) for b in ids[:, 0]] == [b for b, _ in q_ref.blocks()]
    assert np.all(ids[:, 2] == 1.0)         # getParameterBlocks(addr_shift) returned the shifted addresses
    Jm = J.reshape(n, n)
    A0, b0 = q_ref.information()
    assert np.abs(Jm.T @ Jm - A0).max() <= 1e-6 * np.abs(A0).max()
    assert np.abs(Jm.T @ r - b0).max() <= 1e-6 * np.abs(b0).max()


@pytest.fixture(scope="module")
def emu_pair():
    from emu import build_emu
    path = build_emu.build()
    ctx = viwb_lib.Context(0, path)
    yield ctx, build_exe(path, "emu")
    ctx.close()


@pytest.mark.parametrize("cid,with_prior", [(1, False), (4, True)])
def test_shim_optimization_matches_c_abi_emulation(emu_pair, tmp_path, cid, with_prior):
    ctx, exe = emu_pair
    run_case(ctx, exe, tmp_path, cid, with_prior)


@pytest.mark.gpu
@pytest.mark.parametrize("cid,with_prior", [(1, False), (2, True), (4, True)])
def test_shim_optimization_matches_c_abi_gpu(gpu_ctx, tmp_path, cid, with_prior):
    exe = build_exe(viwb_lib.DEFAULT_LIB, "gpu")
    run_case(gpu_ctx, exe, tmp_path, cid, with_prior)

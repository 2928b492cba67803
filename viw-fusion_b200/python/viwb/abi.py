"""ctypes mirror of include/viwb.h (the C ABI of libviwb.so).

Only data-format definitions live here: struct layouts, block tables and a `WindowProblem` holder that
keeps the numpy arrays alive behind a `viwb_problem`.  Python is plumbing for tests and bench.py; the
product is the CUDA library behind the C ABI.
"""
import ctypes as C
import numpy as np

WINDOW_SIZE = 10
NUM_FRAMES = 11
NUM_FIXED_BLOCKS = 32
STATE_FIXED = 207
TANGENT_FIXED = 192
MAX_PRIOR_DIM = 200
VIS_OBS_DOUBLES = 12
IMU_DOUBLES = 287
WHEEL_DOUBLES = 78

BLK_POSE0, BLK_SPEEDBIAS0, BLK_EX_POSE0, BLK_EX_POSE1, BLK_EX_WHEEL = 0, 11, 22, 23, 24
BLK_PLANE_R, BLK_PLANE_Z, BLK_SX, BLK_SY, BLK_SW, BLK_TD, BLK_TD_WHEEL, BLK_LANDMARK0 = 25, 26, 27, 28, 29, 30, 31, 32
BLOCK_PRESENT, BLOCK_CONSTANT = 1, 2
F_PROJ_2F1C, F_PROJ_2F2C, F_PROJ_1F2C, F_IMU, F_WHEEL, F_PLANE = 0, 1, 2, 3, 4, 5
MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1
CONVERGENCE, NO_CONVERGENCE, FAILURE = 0, 1, 2      # enum viwb_termination
LK_USE_INITIAL_FLOW = 4

# parameter block signature of each factor class (global sizes), as in the SizedCostFunction<> templates
FACTOR_RESIDUALS = {F_PROJ_2F1C: 2, F_PROJ_2F2C: 2, F_PROJ_1F2C: 2, F_IMU: 15, F_WHEEL: 6, F_PLANE: 3}
FACTOR_BLOCK_SIZES = {F_PROJ_2F1C: (7, 7, 7, 1, 1), F_PROJ_2F2C: (7, 7, 7, 7, 1, 1), F_PROJ_1F2C: (7, 7, 1, 1),
                      F_IMU: (7, 9, 7, 9), F_WHEEL: (7, 7, 7, 1, 1, 1, 1), F_PLANE: (7, 7, 4, 1)}


def block_size(b):
    return 7 if b < 11 else 9 if b < 22 else 7 if b < 25 else 4 if b == 25 else 1


def block_offset(b):
    return 7 * b if b < 11 else 77 + 9 * (b - 11) if b < 22 else 176 + 7 * (b - 22) if b < 25 else 197 if b == 25 else 201 + (b - 26)


def block_tsize(b):
    return 6 if b < 11 else 9 if b < 22 else 6 if b < 25 else 3 if b == 25 else 1


def block_toffset(b):
    return 6 * b if b < 11 else 66 + 9 * (b - 11) if b < 22 else 165 + 6 * (b - 22) if b < 25 else 183 if b == 25 else 186 + (b - 26)


def block_marg_size(b):
    s = block_size(b)
    return 6 if s == 7 else s


c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Prior(C.Structure):
    _fields_ = [("valid", C.c_int32), ("n", C.c_int32), ("num_blocks", C.c_int32),
                ("block_id", C.c_int32 * NUM_FIXED_BLOCKS), ("block_idx", C.c_int32 * NUM_FIXED_BLOCKS),
                ("x0", c_double_p), ("J", c_double_p), ("r", c_double_p)]


class Globals(C.Structure):
    _fields_ = [("G", C.c_double * 3), ("vis_sqrt_info", C.c_double * 4), ("plane_sqrt_info", C.c_double * 3),
                ("huber_delta", C.c_double)]


class Problem(C.Structure):
    _fields_ = [("frame_count", C.c_int32), ("num_landmarks", C.c_int32),
                ("block_flags", C.c_uint8 * NUM_FIXED_BLOCKS), ("subset_mask", C.c_uint8 * NUM_FIXED_BLOCKS),
                ("num_vis", C.c_int32), ("vis_type", c_int32_p), ("vis_landmark", c_int32_p),
                ("vis_frame_i", c_int32_p), ("vis_frame_j", c_int32_p), ("vis_obs", c_double_p),
                ("num_imu", C.c_int32), ("imu_frame_i", c_int32_p), ("imu_frame_j", c_int32_p), ("imu_data", c_double_p),
                ("num_wheel", C.c_int32), ("wheel_frame_i", c_int32_p), ("wheel_frame_j", c_int32_p), ("wheel_data", c_double_p),
                ("num_plane", C.c_int32), ("plane_frame", c_int32_p),
                ("prior", C.POINTER(Prior)), ("globals", Globals)]


class Options(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("max_solver_time_in_seconds", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32)]


class Summary(C.Structure):
    _fields_ = [("termination_type", C.c_int32), ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_linear_solves", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("final_radius", C.c_double), ("final_mu", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def default_options():
    o = Options()
    o.max_num_iterations = 8
    o.max_solver_time_in_seconds = 0.0
    o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance = 1e-6, 1e-10, 1e-8
    o.initial_trust_region_radius, o.max_trust_region_radius, o.min_trust_region_radius = 1e4, 1e16, 1e-32
    o.min_relative_decrease, o.min_lm_diagonal, o.max_lm_diagonal = 1e-3, 1e-6, 1e32
    o.max_num_consecutive_invalid_steps, o.jacobi_scaling = 5, 1
    return o


def default_globals(g_norm=9.81007, pitch_n=0.01, roll_n=0.01, zpw_n=0.05):
    g = Globals()
    g.G[:] = [0.0, 0.0, g_norm]
    g.vis_sqrt_info[:] = [460.0 / 1.5, 0.0, 0.0, 460.0 / 1.5]
    g.plane_sqrt_info[:] = [1.0 / pitch_n, 1.0 / roll_n, 1.0 / zpw_n]
    g.huber_delta = 1.0
    return g


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int32_p)


class PriorData:
    """Owns the buffers of one viwb_prior (also used as an output buffer for marginalize)."""

    def __init__(self):
        self.x0 = np.zeros(STATE_FIXED)
        self.J = np.zeros(MAX_PRIOR_DIM * MAX_PRIOR_DIM)
        self.r = np.zeros(MAX_PRIOR_DIM)
        self.c = Prior()
        self.c.valid = 0
        self.c.n = 0
        self.c.num_blocks = 0
        self.c.x0, self.c.J, self.c.r = dptr(self.x0), dptr(self.J), dptr(self.r)

    @property
    def valid(self):
        return bool(self.c.valid)

    @property
    def n(self):
        return int(self.c.n)

    def blocks(self):
        return [(int(self.c.block_id[i]), int(self.c.block_idx[i])) for i in range(self.c.num_blocks)]

    def Jmat(self):
        n = self.n
        return self.J[: n * n].reshape(n, n).copy()

    def rvec(self):
        return self.r[: self.n].copy()

    def copy(self):
        p = PriorData()
        p.x0[:] = self.x0
        p.J[:] = self.J
        p.r[:] = self.r
        p.c.valid, p.c.n, p.c.num_blocks = self.c.valid, self.c.n, self.c.num_blocks
        for i in range(NUM_FIXED_BLOCKS):
            p.c.block_id[i] = self.c.block_id[i]
            p.c.block_idx[i] = self.c.block_idx[i]
        return p

    def information(self):
        """(A, b) = (J^T J, J^T r) scattered to the fixed marg-local layout; order-independent comparison."""
        n = self.n
        J, r = self.Jmat(), self.rvec()
        return J.T @ J, J.T @ r

    def to_npz_dict(self, prefix):
        return {prefix + "valid": np.int32(self.c.valid), prefix + "n": np.int32(self.c.n),
                prefix + "block_id": np.array([self.c.block_id[i] for i in range(self.c.num_blocks)], np.int32),
                prefix + "block_idx": np.array([self.c.block_idx[i] for i in range(self.c.num_blocks)], np.int32),
                prefix + "x0": self.x0.copy(), prefix + "J": self.Jmat(), prefix + "r": self.rvec()}

    @staticmethod
    def from_arrays(valid, n, block_id, block_idx, x0, J, r):
        p = PriorData()
        p.c.valid, p.c.n, p.c.num_blocks = int(valid), int(n), len(block_id)
        for i, (b, ix) in enumerate(zip(block_id, block_idx)):
            p.c.block_id[i], p.c.block_idx[i] = int(b), int(ix)
        p.x0[:] = x0
        p.J[: n * n] = np.asarray(J, np.float64).reshape(-1)
        p.r[:n] = r
        return p


class WindowProblem:
    """One Estimator::optimization() problem: numpy tables + the viwb_problem view over them."""

    def __init__(self, frame_count, num_landmarks, block_flags, subset_mask=None,
                 vis_type=(), vis_landmark=(), vis_frame_i=(), vis_frame_j=(), vis_obs=(),
                 imu_frame_i=(), imu_frame_j=(), imu_data=(),
                 wheel_frame_i=(), wheel_frame_j=(), wheel_data=(), plane_frame=(),
                 prior=None, globals_=None):
        self.frame_count = int(frame_count)
        self.num_landmarks = int(num_landmarks)
        self.block_flags = np.asarray(block_flags, np.uint8).copy()
        self.subset_mask = np.zeros(NUM_FIXED_BLOCKS, np.uint8) if subset_mask is None else np.asarray(subset_mask, np.uint8).copy()
        self.vis_type, self.vis_landmark = _i32(vis_type), _i32(vis_landmark)
        self.vis_frame_i, self.vis_frame_j = _i32(vis_frame_i), _i32(vis_frame_j)
        self.vis_obs = _f64(vis_obs).reshape(-1, VIS_OBS_DOUBLES)
        self.imu_frame_i, self.imu_frame_j = _i32(imu_frame_i), _i32(imu_frame_j)
        self.imu_data = _f64(imu_data).reshape(-1, IMU_DOUBLES)
        self.wheel_frame_i, self.wheel_frame_j = _i32(wheel_frame_i), _i32(wheel_frame_j)
        self.wheel_data = _f64(wheel_data).reshape(-1, WHEEL_DOUBLES)
        self.plane_frame = _i32(plane_frame)
        self.prior = prior
        self.globals = globals_ if globals_ is not None else default_globals()
        self._c = None

    @property
    def state_size(self):
        return STATE_FIXED + self.num_landmarks

    def fill(self, p):
        """Fill an existing ctypes Problem struct (used for arrays of problems)."""
        p.frame_count, p.num_landmarks = self.frame_count, self.num_landmarks
        for i in range(NUM_FIXED_BLOCKS):
            p.block_flags[i] = int(self.block_flags[i])
            p.subset_mask[i] = int(self.subset_mask[i])
        p.num_vis = len(self.vis_type)
        p.vis_type, p.vis_landmark = iptr(self.vis_type), iptr(self.vis_landmark)
        p.vis_frame_i, p.vis_frame_j, p.vis_obs = iptr(self.vis_frame_i), iptr(self.vis_frame_j), dptr(self.vis_obs)
        p.num_imu = len(self.imu_frame_i)
        p.imu_frame_i, p.imu_frame_j, p.imu_data = iptr(self.imu_frame_i), iptr(self.imu_frame_j), dptr(self.imu_data)
        p.num_wheel = len(self.wheel_frame_i)
        p.wheel_frame_i, p.wheel_frame_j, p.wheel_data = iptr(self.wheel_frame_i), iptr(self.wheel_frame_j), dptr(self.wheel_data)
        p.num_plane = len(self.plane_frame)
        p.plane_frame = iptr(self.plane_frame)
        p.prior = C.pointer(self.prior.c) if self.prior is not None else None
        p.globals = self.globals
        return p

    @property
    def c(self):
        self._c = self.fill(Problem())
        return self._c

    def algorithmic_bytes(self, iters=8):
        """SURVEY 8(d) B_solve model for this window (FP64 inputs as the reference stores them)."""
        n_vis, n_imu, n_wheel, n_plane = len(self.vis_type), len(self.imu_frame_i), len(self.wheel_frame_i), len(self.plane_frame)
        active = [b for b in range(NUM_FIXED_BLOCKS) if (self.block_flags[b] & BLOCK_PRESENT) and not (self.block_flags[b] & BLOCK_CONSTANT)]
        R = sum(block_tsize(b) for b in active)
        p_glob = sum(block_size(b) for b in range(NUM_FIXED_BLOCKS) if self.block_flags[b] & BLOCK_PRESENT) + self.num_landmarks
        n = self.prior.n if (self.prior is not None and self.prior.valid) else 0
        b_iter = 112 * n_vis + 2296 * n_imu + 624 * n_wheel + 4 * n_plane + 8 * (n * n + 2 * n) + 8 * p_glob + 2 * 8 * (R * R + R) + 8 * p_glob
        n_vis0 = int(np.sum(self.vis_frame_i == 0))
        lm0 = len(set(self.vis_landmark[self.vis_frame_i == 0].tolist()))
        m = 15 + lm0
        nn = max(n, 76)
        b_marg = 112 * n_vis0 + 2296 + 624 * (1 if n_wheel else 0) + 4 * (1 if n_plane else 0) + 8 * (n * n + 2 * n) + 8 * ((m + nn) ** 2 + (m + nn)) + 8 * (nn * nn + nn)
        return iters * b_iter + b_marg

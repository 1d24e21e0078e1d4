"""ctypes mirror of include/bioslam_b200.h (the C ABI) plus a small Python-side graph description builder.

Only plain data lives here: enums, POD structs and `GraphDesc`, the flat time-chain description
(K keyframes x time block + static block + factor slots) that both the CUDA library and the test oracle consume.
"""
import ctypes as C
import numpy as np

# ---- enums (include/bioslam_b200.h) ------------------------------------------------------------------
OK, BAD_ARG, CUDA_ERROR, INDETERMINATE_SYSTEM, LAMBDA_MAXED, NOT_IMPLEMENTED, NCCL_ERROR = range(7)
VAR_POSE3, VAR_VEC3, VAR_BIAS6, VAR_UNIT3 = range(4)
VAL_DIM = (12, 3, 6, 3)
TAN_DIM = (6, 3, 6, 2)
(F_IMU_COMBINED, F_IMU_STATICBIAS, F_ANGVEL, F_HINGE_VEC, F_HINGE_NORM, F_JOINTPOS_VEC, F_JOINTPOS_NORM, F_JOINTVEL_VEC,
 F_SEGLEN_MAG, F_SEGLEN_MAX, F_SEGLEN_MIN, F_SEGLEN_DISCREPANCY, F_ANGLE_AXIS_SEG, F_POSE3_TRANSLATION_PRIOR,
 F_POSE3_COMPASS_PRIOR, F_PRIOR_VEC3, F_PRIOR_BIAS6, F_PRIOR_UNIT3, F_PRIOR_POSE3, F_MAG_POSE3, F_JOINTVEL_NORM, F_NUM_TYPES) = range(22)
AT_K, AT_K1, AT_STATIC = range(3)
MAX_FACTOR_VARS, MAX_FACTOR_PARAMS = 8, 24

# (rows, variable types, per-keyframe constants) per factor type
FACTOR_INFO = {
    F_IMU_COMBINED: (15, (VAR_POSE3, VAR_VEC3, VAR_POSE3, VAR_VEC3, VAR_BIAS6, VAR_BIAS6), 190),
    F_IMU_STATICBIAS: (9, (VAR_POSE3, VAR_VEC3, VAR_POSE3, VAR_VEC3, VAR_BIAS6), 115),
    F_ANGVEL: (3, (VAR_VEC3, VAR_BIAS6), 3),
    F_HINGE_VEC: (3, (VAR_POSE3, VAR_VEC3, VAR_POSE3, VAR_VEC3, VAR_UNIT3), 0),
    F_HINGE_NORM: (1, (VAR_POSE3, VAR_VEC3, VAR_POSE3, VAR_VEC3, VAR_UNIT3), 0),
    F_JOINTPOS_VEC: (3, (VAR_POSE3, VAR_POSE3, VAR_VEC3, VAR_VEC3), 0),
    F_JOINTPOS_NORM: (1, (VAR_POSE3, VAR_POSE3, VAR_VEC3, VAR_VEC3), 0),
    F_JOINTVEL_VEC: (3, (VAR_POSE3, VAR_VEC3, VAR_VEC3, VAR_VEC3, VAR_POSE3, VAR_VEC3, VAR_VEC3, VAR_VEC3), 0),
    F_SEGLEN_MAG: (1, (VAR_VEC3, VAR_VEC3), 0),
    F_SEGLEN_MAX: (1, (VAR_VEC3, VAR_VEC3), 0),
    F_SEGLEN_MIN: (1, (VAR_VEC3, VAR_VEC3), 0),
    F_SEGLEN_DISCREPANCY: (1, (VAR_VEC3, VAR_VEC3, VAR_VEC3, VAR_VEC3), 0),
    F_ANGLE_AXIS_SEG: (1, (VAR_UNIT3, VAR_VEC3, VAR_VEC3), 0),
    F_POSE3_TRANSLATION_PRIOR: (3, (VAR_POSE3,), 0),
    F_POSE3_COMPASS_PRIOR: (1, (VAR_POSE3,), 0),
    F_PRIOR_VEC3: (3, (VAR_VEC3,), 0),
    F_PRIOR_BIAS6: (6, (VAR_BIAS6,), 0),
    F_PRIOR_UNIT3: (2, (VAR_UNIT3,), 0),
    F_PRIOR_POSE3: (6, (VAR_POSE3,), 0),
    F_MAG_POSE3: (3, (VAR_POSE3,), 3),
    F_JOINTVEL_NORM: (1, (VAR_POSE3, VAR_VEC3, VAR_VEC3, VAR_VEC3, VAR_POSE3, VAR_VEC3, VAR_VEC3, VAR_VEC3), 0),
}


class FactorSlot(C.Structure):
    _fields_ = [("type", C.c_int32), ("k_begin", C.c_int32), ("k_end", C.c_int32), ("nvars", C.c_int32),
                ("var_kind", C.c_int32 * MAX_FACTOR_VARS), ("var_index", C.c_int32 * MAX_FACTOR_VARS),
                ("const_offset", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * MAX_FACTOR_PARAMS)]


class ProblemDesc(C.Structure):
    _fields_ = [("K", C.c_int32), ("n_time_vars", C.c_int32), ("time_var_type", C.POINTER(C.c_int32)),
                ("n_static_vars", C.c_int32), ("static_var_type", C.POINTER(C.c_int32)),
                ("n_slots", C.c_int32), ("slots", C.POINTER(FactorSlot)),
                ("const_stride", C.c_int32), ("consts", C.POINTER(C.c_double))]


class LmParams(C.Structure):
    _fields_ = [("lambda_initial", C.c_double), ("lambda_factor", C.c_double), ("lambda_lower_bound", C.c_double),
                ("lambda_upper_bound", C.c_double), ("relative_error_tol", C.c_double), ("absolute_error_tol", C.c_double),
                ("min_model_fidelity", C.c_double), ("gauss_newton", C.c_int32), ("reserved", C.c_int32)]

    @staticmethod
    def lower_body():
        """reference include/lowerBodyPoseEstimator.h:82-89, src/lowerBodyPoseEstimator.cpp:644-647"""
        return LmParams(1.0e-5, 10.0, 1.0e-10, 1.0e25, 1.0e-20, 1.0e-20, 1.0e-3, 0, 0)

    @staticmethod
    def single_imu():
        """reference src/imuPoseEstimator.cpp:324-325 (lambda bounds 1e-8..1e10), include/imuPoseEstimator.h:49-52"""
        return LmParams(1.0e-5, 10.0, 1.0e-8, 1.0e10, 1.0e-8, 1.0e-10, 1.0e-3, 0, 0)


class LmState(C.Structure):
    _fields_ = [("error", C.c_double), ("lambda_", C.c_double), ("iterations", C.c_int32), ("inner_trials", C.c_int32),
                ("total_trials", C.c_int32), ("status", C.c_int32)]


class OuterParams(C.Structure):
    _fields_ = [("abs_err_decrease_limit", C.c_double), ("rel_err_decrease_limit", C.c_double), ("abs_err_limit", C.c_double),
                ("max_iterations", C.c_int32), ("converge_at_consec_small_iter", C.c_int32), ("single_imu_rule", C.c_int32),
                ("max_history", C.c_int32)]

    @staticmethod
    def lower_body(max_iterations=10000, max_history=20000):
        return OuterParams(1.0e-6, 1.0e-5, 1.0e-5, max_iterations, 6, 0, max_history)

    @staticmethod
    def single_imu(max_iterations=5000, max_history=20000):
        return OuterParams(1.0e-10, 1.0e-8, 1.0e-12, max_iterations, 1, 1, max_history)


class OptimizeResult(C.Structure):
    _fields_ = [("initial_error", C.c_double), ("final_error", C.c_double), ("final_lambda", C.c_double),
                ("iterations", C.c_int32), ("outer_calls", C.c_int32), ("total_trials", C.c_int32), ("n_history", C.c_int32),
                ("status", C.c_int32), ("reserved", C.c_int32), ("device_ms", C.c_double)]


class ImuNoise(C.Structure):
    _fields_ = [("accel_noise_sigma", C.c_double), ("accel_bias_rw_sigma", C.c_double), ("gyro_noise_sigma", C.c_double),
                ("gyro_bias_rw_sigma", C.c_double), ("preint_bias_sigma", C.c_double), ("integration_err_sigma", C.c_double)]

    @staticmethod
    def opal_v2():
        """reference include/imu/imuNoiseModelHandler.h:28-35"""
        return ImuNoise(0.04903325, 0.005905, 0.00015707963267949, 1.93925e-5, 3.162e-3, 1.0e-4)


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class JointAngleDesc(C.Structure):
    """b200_joint_angle_desc (include/bioslam_b200.h)"""
    _fields_ = [("pose_var", C.c_int32 * 5), ("static_var", C.c_int32 * 14)]


JOINT_ANGLE_NAMES = [j + "_" + a for j in ("rknee", "lknee", "rhip", "lhip") for a in ("flexEx", "abAd", "intExtRot")]


class GraphDesc:
    """Python-side holder of a b200_problem_desc (keeps the numpy buffers alive)."""

    def __init__(self, K, time_var_types, static_var_types=(), const_stride=0):
        self.K = int(K)
        self.time_var_types = list(time_var_types)
        self.static_var_types = list(static_var_types)
        self.const_stride = int(const_stride)
        self.consts = np.zeros((self.K, max(self.const_stride, 1)), dtype=np.float64)
        self.slots = []

    # dims
    @property
    def time_value_dim(self): return sum(VAL_DIM[t] for t in self.time_var_types)
    @property
    def static_value_dim(self): return sum(VAL_DIM[t] for t in self.static_var_types)
    @property
    def nt(self): return sum(TAN_DIM[t] for t in self.time_var_types)
    @property
    def ns(self): return sum(TAN_DIM[t] for t in self.static_var_types)

    def time_value_offsets(self):
        return np.concatenate([[0], np.cumsum([VAL_DIM[t] for t in self.time_var_types])]).astype(int)

    def static_value_offsets(self):
        return np.concatenate([[0], np.cumsum([VAL_DIM[t] for t in self.static_var_types])]).astype(int)

    def time_tangent_offsets(self):
        return np.concatenate([[0], np.cumsum([TAN_DIM[t] for t in self.time_var_types])]).astype(int)

    def static_tangent_offsets(self):
        return np.concatenate([[0], np.cumsum([TAN_DIM[t] for t in self.static_var_types])]).astype(int)

    def add_slot(self, ftype, vars_, params=(), k_begin=0, k_end=None, const_offset=-1):
        """vars_: list of (kind, index).  A slot with only static variables must keep k_begin=0,k_end=1."""
        rows, vtypes, nconst = FACTOR_INFO[ftype]
        assert len(vars_) == len(vtypes), (ftype, vars_)
        all_static = all(k == AT_STATIC for k, _ in vars_)
        if k_end is None:
            k_end = 1 if all_static else (self.K - 1 if any(k == AT_K1 for k, _ in vars_) else self.K)
        s = FactorSlot()
        s.type, s.k_begin, s.k_end, s.nvars = ftype, k_begin, k_end, len(vars_)
        for i, (kind, idx) in enumerate(vars_):
            tab = self.static_var_types if kind == AT_STATIC else self.time_var_types
            assert tab[idx] == vtypes[i], f"factor {ftype} var {i}: table type {tab[idx]} != {vtypes[i]}"
            s.var_kind[i], s.var_index[i] = kind, idx
        s.const_offset = const_offset
        assert (nconst == 0) == (const_offset < 0)
        for i, v in enumerate(params):
            s.params[i] = float(v)
        self.slots.append(s)
        return s

    def c_desc(self):
        self._tv = (C.c_int32 * len(self.time_var_types))(*self.time_var_types)
        self._sv = (C.c_int32 * max(len(self.static_var_types), 1))(*self.static_var_types)
        self._sl = (FactorSlot * len(self.slots))(*self.slots)
        self.consts = np.ascontiguousarray(self.consts, dtype=np.float64)
        d = ProblemDesc()
        d.K, d.n_time_vars, d.time_var_type = self.K, len(self.time_var_types), self._tv
        d.n_static_vars, d.static_var_type = len(self.static_var_types), self._sv
        d.n_slots, d.slots = len(self.slots), self._sl
        d.const_stride, d.consts = self.const_stride, dptr(self.consts)
        return d

    @staticmethod
    def from_c(desc):
        """deep-copy a ProblemDesc produced by the C++ host layer into a GraphDesc"""
        g = GraphDesc(desc.K, [desc.time_var_type[i] for i in range(desc.n_time_vars)],
                      [desc.static_var_type[i] for i in range(desc.n_static_vars)], desc.const_stride)
        if desc.const_stride > 0:
            g.consts = np.ctypeslib.as_array(desc.consts, shape=(desc.K, desc.const_stride)).copy()
        for i in range(desc.n_slots):
            s = FactorSlot()
            C.memmove(C.byref(s), C.byref(desc.slots[i]), C.sizeof(FactorSlot))
            g.slots.append(s)
        return g

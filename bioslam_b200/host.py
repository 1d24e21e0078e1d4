"""ctypes access to the C++ host layer (bioslam_b200/host: imu, imuPoseEstimator, lowerBodyPoseEstimator mirrors of the
reference classes) that lives inside libbioslam_b200.so."""
import ctypes as C
import numpy as np
from . import abi, lib

D = abi.dptr


class HostError(RuntimeError):
    pass


def _L(so_path=None):
    L = lib.load(so_path)
    if not getattr(L, "_host_ready", False):
        L.b200h_last_error.restype = C.c_char_p
        for f in ("b200h_imu_create", "b200h_ipe_create", "b200h_lbpe_create"):
            getattr(L, f).restype = C.c_void_p
        L.b200h_imu_create.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p]
        L.b200h_ipe_create.argtypes = [C.c_void_p, C.c_char_p]
        L.b200h_lbpe_create.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
        L.b200h_lbpe_desc.restype = C.POINTER(abi.ProblemDesc)
        for f in ("b200h_imu_destroy", "b200h_ipe_destroy", "b200h_lbpe_destroy", "b200h_ipe_setup", "b200h_ipe_set_optimized_to_initial",
                  "b200h_ipe_fast_optimize", "b200h_ipe_n_keyframes", "b200h_lbpe_setup", "b200h_lbpe_fast_optimize", "b200h_lbpe_set_joint_angles", "b200h_lbpe_desc",
                  "b200h_lbpe_post_hoc_axes_check", "b200h_lbpe_set_members_from_values", "b200h_lbpe_adjust_legs", "b200h_lbpe_n_restarts"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("b200h_ipe_set", "b200h_lbpe_set"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
        for f in ("b200h_ipe_get", "b200h_lbpe_get"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
        L.b200h_lbpe_hip_ie_slopes.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.b200h_ekf_forward_backward.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_int, C.POINTER(C.c_double)]
        L.b200h_lbpe_values.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.b200h_lbpe_set_values.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L._host_ready = True
    return L


def _ck(L, rc):
    if rc != 0:
        raise HostError(L.b200h_last_error().decode())


def ekf_forward_backward(gyro, acc, dt, passes=3, so_path=None):
    """forward/backward orientation EKF (reference src/mathutils.cpp:489-610): quaternions (w, x, y, z) of R[N->B], [n][4]"""
    L = _L(so_path)
    gyro = np.ascontiguousarray(gyro, dtype=np.float64); acc = np.ascontiguousarray(acc, dtype=np.float64)
    q = np.zeros((gyro.shape[0], 4))
    _ck(L, L.b200h_ekf_forward_backward(gyro.shape[0], D(gyro), D(acc), float(dt), int(passes), D(q)))
    return q


class Imu:
    def __init__(self, acc, gyro, t, label="", mag=None, so_path=None):
        self.L = _L(so_path)
        acc = np.ascontiguousarray(acc, dtype=np.float64); gyro = np.ascontiguousarray(gyro, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        m = None if mag is None else D(np.ascontiguousarray(mag, dtype=np.float64))
        self.h = C.c_void_p(self.L.b200h_imu_create(acc.shape[0], D(acc), D(gyro), m, D(t), label.encode()))

    def __del__(self):
        try:
            self.L.b200h_imu_destroy(self.h)
        except Exception:
            pass


class _Settable:
    def set(self, **kw):
        for k, v in kw.items():
            a = np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel().copy()
            _ck(self.L, self._set(self.h, k.encode(), D(a), a.size))
        return self

    def get(self, name, cap=1 << 24):
        out = np.zeros(cap)
        n = self._get(self.h, name.encode(), D(out), cap)
        if n < 0:
            raise HostError(self.L.b200h_last_error().decode())
        return out[:n].copy()


class ImuPoseEstimator(_Settable):
    """mirror of bioslam's imuPoseEstimator (reference include/imuPoseEstimator.h)"""

    def __init__(self, imu: Imu, id_, so_path=None):
        self.L = _L(so_path); self.imu = imu
        self.h = C.c_void_p(self.L.b200h_ipe_create(imu.h, id_.encode()))
        self._set, self._get = self.L.b200h_ipe_set, self.L.b200h_ipe_get

    def setImuBiasModelMode(self, mode): return self.set(imuBiasModelMode=mode)
    def setup(self): _ck(self.L, self.L.b200h_ipe_setup(self.h)); return self
    def setOptimizedValuesToInitialValues(self): _ck(self.L, self.L.b200h_ipe_set_optimized_to_initial(self.h)); return self
    def fastOptimize(self): _ck(self.L, self.L.b200h_ipe_fast_optimize(self.h)); return self
    @property
    def nKeyframes(self): return self.L.b200h_ipe_n_keyframes(self.h)

    def __del__(self):
        try:
            self.L.b200h_ipe_destroy(self.h)
        except Exception:
            pass


class LowerBodyPoseEstimator(_Settable):
    """mirror of bioslam's lowerBodyPoseEstimator (reference include/lowerBodyPoseEstimator.h); IMU order: sacrum, rthigh,
    rshank, rfoot, lthigh, lshank, lfoot"""

    def __init__(self, ipes, id_, so_path=None):
        assert len(ipes) == 7
        self.L = _L(so_path); self.ipes = ipes
        arr = (C.c_void_p * 7)(*[p.h for p in ipes])
        self.h = C.c_void_p(self.L.b200h_lbpe_create(arr, id_.encode()))
        self._set, self._get = self.L.b200h_lbpe_set, self.L.b200h_lbpe_get

    def setup(self): _ck(self.L, self.L.b200h_lbpe_setup(self.h)); return self
    def fastOptimize(self): _ck(self.L, self.L.b200h_lbpe_fast_optimize(self.h)); return self
    def setDerivedJointAnglesFromEstimatedStates(self): _ck(self.L, self.L.b200h_lbpe_set_joint_angles(self.h)); return self
    def setMemberEstimatedStatesFromValues(self): _ck(self.L, self.L.b200h_lbpe_set_members_from_values(self.h)); return self

    def computeHipIeRotAngleRegressionSlopesFromValues(self):
        """reference src/lowerBodyPoseEstimator.cpp:1507-1558 -> (right, left) slope in rad/s"""
        out = np.zeros(2)
        _ck(self.L, self.L.b200h_lbpe_hip_ie_slopes(self.h, D(out)))
        return float(out[0]), float(out[1])

    def adjustLegsForCorrectedHipIeAngles(self): _ck(self.L, self.L.b200h_lbpe_adjust_legs(self.h)); return self
    @property
    def nRestarts(self): return self.L.b200h_lbpe_n_restarts(self.h)

    def postHocAxesCheck(self):
        """reference src/lowerBodyPoseEstimator.cpp:711-797; True if a knee shank axis was flipped"""
        rc = self.L.b200h_lbpe_post_hoc_axes_check(self.h)
        if rc < 0:
            raise HostError(self.L.b200h_last_error().decode())
        return bool(rc)

    def graph(self):
        """deep copy of the flat description setup() produced + its current values"""
        d = self.L.b200h_lbpe_desc(self.h)
        if not d:
            raise HostError("setup() must be called first")
        g = abi.GraphDesc.from_c(d.contents)
        tv = np.zeros((g.K, g.time_value_dim)); sv = np.zeros(g.static_value_dim)
        _ck(self.L, self.L.b200h_lbpe_values(self.h, D(tv), D(sv)))
        return g, tv, sv

    def set_values(self, tv, sv):
        tv = np.ascontiguousarray(tv, dtype=np.float64); sv = np.ascontiguousarray(sv, dtype=np.float64)
        _ck(self.L, self.L.b200h_lbpe_set_values(self.h, D(tv), D(sv)))

    def __del__(self):
        try:
            self.L.b200h_lbpe_destroy(self.h)
        except Exception:
            pass


def lower_body_from_trial(trial, n_per=10, so_path=None, example_priors=True):
    """the call sequence of reference examples/exampleLowerBodyEstimation.cpp:16-153 on a synthetic trial"""
    ipes = []
    for i in range(7):
        m = Imu(trial.acc[i], trial.gyro[i], trial.time, label=str(i), so_path=so_path)
        p = ImuPoseEstimator(m, f"imu{i}", so_path=so_path)
        p.set(m_numMeasToPreint=n_per, m_globalAcc=[0.0, 0.0, -9.81])
        p.setImuBiasModelMode(1)
        if i == 0 and example_priors:
            p.set(m_useCompassPrior=1, m_refVecLocal=[0, 1, 0], m_refVecNav=[1, 0, 0], m_usePositionPrior=1, m_useVelocityPrior=1, m_useImuBiasPrior=0)
        else:
            p.set(m_usePositionPrior=0, m_useVelocityPrior=0, m_useImuBiasPrior=0)
        p.setup().setOptimizedValuesToInitialValues()
        ipes.append(p)
    return LowerBodyPoseEstimator(ipes, "example", so_path=so_path)

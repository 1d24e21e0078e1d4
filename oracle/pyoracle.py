"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/liboracle.so (build: `make -C oracle`)."""
import ctypes as C
import os
import subprocess
import numpy as np
from bioslam_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.ProblemDesc)]
        L.orc_error.restype = C.c_double
        for f in ("orc_destroy", "orc_set_threads", "orc_set_values", "orc_get_values", "orc_linearize", "orc_get_gradient",
                  "orc_get_hessian_blocks", "orc_get_static_hessian", "orc_retract", "orc_lm_begin", "orc_set_solver"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


D = abi.dptr


class OracleProblem:
    def __init__(self, g: abi.GraphDesc, threads=1):
        self.g = g
        self.L = lib()
        d = g.c_desc()
        self.h = C.c_void_p(self.L.orc_create(C.byref(d)))
        self.K, self.nt, self.ns = g.K, g.nt, g.ns
        self.tvd, self.svd = g.time_value_dim, g.static_value_dim
        self.L.orc_set_threads(self.h, C.c_int(threads))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set_values(self, tv, sv):
        tv = np.ascontiguousarray(tv, dtype=np.float64); sv = np.ascontiguousarray(sv, dtype=np.float64)
        assert tv.shape == (self.K, self.tvd) and sv.size == self.svd
        self.L.orc_set_values(self.h, D(tv), D(sv))

    def get_values(self):
        tv = np.zeros((self.K, self.tvd)); sv = np.zeros(max(self.svd, 1))
        self.L.orc_get_values(self.h, D(tv), D(sv))
        return tv, sv[:self.svd]

    def set_solver(self, fast=True, chunks=16):
        """fast: the structure-aware solver of oracle/fastsolve.hpp (results depend on `chunks`, not on the thread count)"""
        self.L.orc_set_solver(self.h, C.c_int(1 if fast else 0), C.c_int(chunks))

    def error(self):
        return float(self.L.orc_error(self.h))

    def linearize(self):
        self.L.orc_linearize(self.h)

    def gradient(self):
        g = np.zeros((self.K, self.nt)); gs = np.zeros(max(self.ns, 1))
        self.L.orc_get_gradient(self.h, D(g), D(gs))
        return g, gs[:self.ns]

    def hessian_blocks(self, k):
        Dm = np.zeros((self.nt, self.nt)); O = np.zeros((self.nt, self.nt)); A = np.zeros((self.nt, max(self.ns, 1)))
        A = np.zeros(self.nt * max(self.ns, 1))
        self.L.orc_get_hessian_blocks(self.h, C.c_int(k), D(Dm), D(O), D(A))
        return Dm, O, A[:self.nt * self.ns].reshape(self.nt, self.ns)

    def static_hessian(self):
        S = np.zeros(max(self.ns * self.ns, 1))
        self.L.orc_get_static_hessian(self.h, D(S))
        return S[:self.ns * self.ns].reshape(self.ns, self.ns)

    def solve(self, lam):
        dt = np.zeros((self.K, self.nt)); ds = np.zeros(max(self.ns, 1))
        st = self.L.orc_solve(self.h, C.c_double(lam), D(dt), D(ds))
        return st, dt, ds[:self.ns]

    def retract(self, dt, ds):
        dt = np.ascontiguousarray(dt); ds = np.ascontiguousarray(np.concatenate([ds, [0.0]]))
        self.L.orc_retract(self.h, D(dt), D(ds))

    def lm_begin(self, params):
        e = C.c_double()
        self.L.orc_lm_begin(self.h, C.byref(params), C.byref(e))
        return e.value

    def lm_iterate(self):
        st = abi.LmState()
        self.L.orc_lm_iterate(self.h, C.byref(st))
        return st

    def optimize(self, lm, outer):
        hist = np.zeros(max(outer.max_history, 1))
        res = abi.OptimizeResult()
        self.L.orc_optimize(self.h, C.byref(lm), C.byref(outer), D(hist), C.byref(res))
        return res, hist[:res.n_history].copy()


def factor_eval(ftype, params, consts, xcat, whiten=False, jac=True):
    L = lib()
    rows, cols = L.orc_factor_rows(ftype), L.orc_factor_cols(ftype)
    p = np.zeros(abi.MAX_FACTOR_PARAMS); p[:len(params)] = params
    c = np.ascontiguousarray(consts if consts is not None and len(consts) else np.zeros(1), dtype=np.float64)
    x = np.ascontiguousarray(xcat, dtype=np.float64)
    r = np.zeros(rows); J = np.zeros((rows, cols))
    L.orc_factor_eval(C.c_int(ftype), D(p), D(c), D(x), D(r), D(J) if jac else None, C.c_int(1 if whiten else 0))
    return (r, J) if jac else r


def retract_var(vtype, x, d):
    out = np.zeros(abi.VAL_DIM[vtype])
    lib().orc_retract_var(C.c_int(vtype), D(np.ascontiguousarray(x, dtype=np.float64)), D(np.ascontiguousarray(d, dtype=np.float64)), D(out))
    return out


def so3_expmap(w):
    R = np.zeros(9); lib().orc_so3_expmap(D(np.ascontiguousarray(w, dtype=np.float64)), D(R)); return R.reshape(3, 3)


def so3_logmap(R):
    w = np.zeros(3); lib().orc_so3_logmap(D(np.ascontiguousarray(R, dtype=np.float64).ravel()), D(w)); return w


def preintegrate(acc, gyro, dt, n_per, bias_hat, noise=None, combined=True):
    """acc, gyro: [n_imu][n_samples][3]; returns [n_imu][n_intervals][190|115]"""
    acc = np.ascontiguousarray(acc, dtype=np.float64); gyro = np.ascontiguousarray(gyro, dtype=np.float64)
    n_imu, n_samples, _ = acc.shape
    K = (n_samples + n_per - 1) // n_per
    n_int = K - 1
    rec = 190 if combined else 115
    out = np.zeros((n_imu, n_int, rec))
    noise = noise or abi.ImuNoise.opal_v2()
    bh = np.ascontiguousarray(bias_hat, dtype=np.float64)
    st = lib().orc_preintegrate(C.c_int(n_imu), C.c_int(n_samples), C.c_int(n_per), C.c_int(n_int), D(acc), D(gyro), C.c_double(dt),
                                D(bh), C.byref(noise), C.c_int(1 if combined else 0), D(out))
    assert st == 0, st
    return out


# ---- derived joint angles (oracle/jointangles.hpp)
def R_imu_to_segment(right_axis, to_prox, to_dist, prox_is_master=True):
    R = np.zeros(9)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (right_axis, to_prox, to_dist)]
    lib().orc_R_imu_to_segment(D(a[0]), D(a[1]), D(a[2]), C.c_int(1 if prox_is_master else 0), D(R))
    return R.reshape(3, 3)


def R_sacrum_to_pelvis(up, to_rhip, to_lhip, hip_is_master=True):
    R = np.zeros(9)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (up, to_rhip, to_lhip)]
    lib().orc_R_sacrum_to_pelvis(D(a[0]), D(a[1]), D(a[2]), C.c_int(1 if hip_is_master else 0), D(R))
    return R.reshape(3, 3)


def R_segment_to_N(R_imu_to_N, R_imu_to_seg):
    R = np.zeros(9)
    lib().orc_R_segment_to_N(D(np.ascontiguousarray(R_imu_to_N, dtype=np.float64).ravel()), D(np.ascontiguousarray(R_imu_to_seg, dtype=np.float64).ravel()), D(R))
    return R.reshape(3, 3)


def jcs_angles(R_prox_to_N, R_dist_to_N, clinical=False, right_leg=True):
    """[flexEx, abAd, intExtRot] (rad): bioutils::consistent3DofJcsAngles / clinical3DofJcsAngles"""
    ang = np.zeros(3)
    lib().orc_jcs_angles(D(np.ascontiguousarray(R_prox_to_N, dtype=np.float64).ravel()), D(np.ascontiguousarray(R_dist_to_N, dtype=np.float64).ravel()),
                         C.c_int(1 if clinical else 0), C.c_int(1 if right_leg else 0), D(ang))
    return ang


def lower_body_joint_angles(g, tv, sv, desc):
    """[K][12] from value arrays in GraphDesc layout (same descriptor as b200_lower_body_joint_angles)"""
    toff, soff = g.time_value_offsets(), g.static_value_offsets()
    K = tv.shape[0]
    R = np.ascontiguousarray(np.stack([tv[:, toff[desc.pose_var[i]]:toff[desc.pose_var[i]] + 9] for i in range(5)]), dtype=np.float64)
    st = np.ascontiguousarray(np.concatenate([sv[soff[desc.static_var[i]]:soff[desc.static_var[i]] + 3] for i in range(14)]), dtype=np.float64)
    out = np.zeros((K, 12))
    lib().orc_lower_body_joint_angles(C.c_int(K), D(R), D(st), D(out))
    return out

"""ctypes binding of libbioslam_b200.so (the CUDA library behind include/bioslam_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os
import numpy as np
from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.path.join(_HERE, "libbioslam_b200.so")
_libs = {}

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p)


class B200Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"bioslam_b200 status {status}: {msg}")
        self.status = status


def load(path=None):
    path = path or DEFAULT_SO
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "bioslam_b200 has no CPU fallback.")
    L = C.CDLL(path)
    L.b200_last_error.restype = C.c_char_p
    L.b200_problem_create.argtypes = [C.POINTER(abi.ProblemDesc), C.c_int32, C.POINTER(C.c_void_p)]
    L.b200_problem_destroy.argtypes = [C.c_void_p]
    L.b200_problem_destroy.restype = None
    L.b200_launch_count.restype = C.c_int64
    L.b200_launch_count.argtypes = [C.c_void_p]
    for f in ("b200_time_value_dim", "b200_static_value_dim", "b200_time_tangent_dim", "b200_static_tangent_dim", "b200_get_chunks", "b200_get_groups"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.b200_comm_set.argtypes = [C.c_void_p, C.c_int32, C.c_int32, ALLGATHER_FN, ALLREDUCE_FN, C.c_void_p]
    L.b200_spec_set.argtypes = [C.c_void_p, C.c_int32, C.c_int32, ALLGATHER_FN, C.c_void_p]
    L.b200_get_levels.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32]
    L.b200_debug_get_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.b200_debug_local_elim.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
    L.b200_nccl_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
    L.b200_comm_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]
    L.b200_host_sync_count.argtypes = [C.c_void_p]
    L.b200_host_sync_count.restype = C.c_int64
    L.b200_lower_body_joint_angles.argtypes = [C.c_void_p, C.POINTER(abi.JointAngleDesc), C.c_void_p]
    _libs[path] = L
    return L


D = abi.dptr


class Problem:
    """One device-resident factor-graph problem == gtsam::LevenbergMarquardtOptimizer(graph, values, params) on a B200."""

    def __init__(self, g: abi.GraphDesc, device=0, so_path=None, chunks=None, groups=None):
        self.L = load(so_path)
        self.g = g
        self.h = C.c_void_p()
        d = g.c_desc()
        self._ck(self.L.b200_problem_create(C.byref(d), C.c_int32(device), C.byref(self.h)))
        self.K, self.nt, self.ns = g.K, g.nt, g.ns
        self.tvd, self.svd = g.time_value_dim, g.static_value_dim
        if chunks is not None and groups is None:
            self.set_chunks(chunks)
        elif chunks is not None or groups is not None:
            self.set_partition(chunks or 0, groups or 0)

    def _ck(self, st, ok=(abi.OK,)):
        if st not in ok:
            raise B200Error(st, self.L.b200_last_error().decode())
        return st

    def close(self):
        if self.h:
            self.L.b200_problem_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_chunks(self, n):
        self._ck(self.L.b200_set_chunks(self.h, C.c_int32(n)))

    def set_partition(self, chunks, groups):
        self._ck(self.L.b200_set_partition(self.h, C.c_int32(chunks), C.c_int32(groups)))

    @property
    def chunks(self):
        return self.L.b200_get_chunks(self.h)

    @property
    def groups(self):
        return self.L.b200_get_groups(self.h)

    def last_trial_scalars(self):
        out = (C.c_double * 4)()
        self._ck(self.L.b200_debug_get_scalars(self.h, out))
        return list(out)

    def joint_angles(self, desc):
        """[K][12] clinical hip / knee angles (rad) of the current values, computed on the device (abi.JOINT_ANGLE_NAMES)"""
        out = np.zeros((self.K, 12))
        self._ck(self.L.b200_lower_body_joint_angles(self.h, C.byref(desc), D(out)))
        return out

    @property
    def levels(self):
        """chunk counts of the nested partition, level 0 first ([] = plain sequential sweep)"""
        sizes = (C.c_int32 * 16)()
        n = self.L.b200_get_levels(self.h, sizes, C.c_int32(16))
        return [int(sizes[i]) for i in range(min(n, 16))]

    def set_values(self, tv, sv):
        tv = np.ascontiguousarray(tv, dtype=np.float64)
        sv = np.ascontiguousarray(sv if self.svd else np.zeros(1), dtype=np.float64)
        assert tv.shape == (self.K, self.tvd) and (sv.size == self.svd or not self.svd)
        self._ck(self.L.b200_set_values(self.h, D(tv), D(sv)))

    def get_values(self, out=None):
        """out = (tv, sv): caller-owned C-contiguous float64 buffers of shapes (K, tvd) / (>= max(svd, 1),), e.g. pinned host memory
        (the library copies straight into them: cudaMemcpyAsync device -> host on the problem's stream)"""
        if out is None:
            tv = np.zeros((self.K, self.tvd)); sv = np.zeros(max(self.svd, 1))
        else:
            tv, sv = out
            assert tv.dtype == np.float64 and sv.dtype == np.float64 and tv.flags.c_contiguous and tv.shape == (self.K, self.tvd) and sv.size >= max(self.svd, 1)
        self._ck(self.L.b200_get_values(self.h, D(tv), D(sv)))
        return tv, sv[:self.svd]

    def error(self):
        e = C.c_double()
        self._ck(self.L.b200_error(self.h, C.byref(e)))
        return e.value

    def lm_begin(self, params):
        e = C.c_double()
        self._ck(self.L.b200_lm_begin(self.h, C.byref(params), C.byref(e)))
        return e.value

    def lm_iterate(self):
        st = abi.LmState()
        self._ck(self.L.b200_lm_iterate(self.h, C.byref(st)), ok=(abi.OK, abi.LAMBDA_MAXED, abi.INDETERMINATE_SYSTEM))
        return st

    def optimize(self, lm, outer):
        hist = np.zeros(max(outer.max_history, 1))
        res = abi.OptimizeResult()
        self._ck(self.L.b200_optimize(self.h, C.byref(lm), C.byref(outer), D(hist), C.byref(res)),
                 ok=(abi.OK, abi.LAMBDA_MAXED, abi.INDETERMINATE_SYSTEM))
        return res, hist[:res.n_history].copy()

    def set_option(self, name, value):
        self._ck(self.L.b200_set_option(self.h, name.encode(), C.c_int32(int(value))))

    def host_sync_count(self):
        return int(self.L.b200_host_sync_count(self.h))

    def launch_count(self):
        return int(self.L.b200_launch_count(self.h))

    def phase_ms(self, reset=False):
        out = (C.c_double * 12)()
        self._ck(self.L.b200_phase_ms(self.h, out, C.c_int32(1 if reset else 0)))
        return list(out)

    def timer_begin(self):
        self._ck(self.L.b200_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_double()
        self._ck(self.L.b200_timer_end(self.h, C.byref(ms)))
        return ms.value

    # ---- parity/debug accessors
    def linearize(self):
        self._ck(self.L.b200_debug_linearize(self.h))

    def gradient(self):
        g = np.zeros((self.K, self.nt)); gs = np.zeros(max(self.ns, 1))
        self._ck(self.L.b200_debug_get_gradient(self.h, D(g), D(gs)))
        return g, gs[:self.ns]

    def hessian_blocks(self, k):
        Dm = np.zeros((self.nt, self.nt)); O = np.zeros((self.nt, self.nt)); A = np.zeros(self.nt * max(self.ns, 1))
        self._ck(self.L.b200_debug_get_hessian_blocks(self.h, C.c_int32(k), D(Dm), D(O), D(A)))
        return Dm, O, A[:self.nt * self.ns].reshape(self.nt, self.ns)

    def static_hessian(self):
        S = np.zeros(max(self.ns * self.ns, 1))
        self._ck(self.L.b200_debug_get_static_hessian(self.h, D(S)))
        return S[:self.ns * self.ns].reshape(self.ns, self.ns)

    def local_elim(self, k, lam):
        """(chain_ext, Dc, Ac, Sc) of keyframe k: the blocks left after the keyframe-local dofs are eliminated at damping lam
        (None if this problem has no pre-elimination)"""
        dims = (C.c_int32 * 4)()
        self._ck(self.L.b200_debug_local_elim(self.h, C.c_int32(k), C.c_double(lam), dims, None, None, None, None))
        nl, nc, ns1, used = list(dims)
        if not used:
            return None
        ext = (C.c_int32 * nc)(); Dc = np.zeros((nc, nc)); Ac = np.zeros((ns1, nc)); Sc = np.zeros((ns1, ns1))
        self._ck(self.L.b200_debug_local_elim(self.h, C.c_int32(k), C.c_double(lam), dims, ext, D(Dc), D(Ac), D(Sc)))
        return np.array(list(ext)), Dc, Ac, Sc

    def solve(self, lam):
        dt = np.zeros((self.K, self.nt)); ds = np.zeros(max(self.ns, 1))
        st = self.L.b200_debug_solve(self.h, C.c_double(lam), D(dt), D(ds))
        self._ck(st, ok=(abi.OK, abi.INDETERMINATE_SYSTEM))
        return st, dt, ds[:self.ns]

    def comm_set(self, world, rank, allgather, allreduce):
        self._ag = ALLGATHER_FN(allgather); self._ar = ALLREDUCE_FN(allreduce)
        self._ck(self.L.b200_comm_set(self.h, C.c_int32(world), C.c_int32(rank), self._ag, self._ar, None))

    def nccl_init(self, n_ranks, global_rank, unique_id, n_groups=1):
        """collectives inside the library (NCCL resolved with dlopen): unique_id = the 128 bytes rank 0 got from nccl_unique_id()"""
        self._ck(self.L.b200_nccl_init(self.h, C.c_int32(n_ranks), C.c_int32(global_rank), C.c_char_p(bytes(unique_id)), C.c_int32(n_groups)))

    def comm_ms(self, reset=False):
        ms = C.c_double(); calls = C.c_int64()
        self._ck(self.L.b200_comm_ms(self.h, C.byref(ms), C.byref(calls), C.c_int32(1 if reset else 0)))
        return ms.value, int(calls.value)

    def spec_set(self, n_groups, group, allgather):
        """lambda speculation: this process is in group `group` of `n_groups` (call after comm_set)"""
        self._spec_ag = ALLGATHER_FN(allgather)
        self._ck(self.L.b200_spec_set(self.h, C.c_int32(n_groups), C.c_int32(group), self._spec_ag, None))


def nccl_unique_id(so_path=None):
    L = load(so_path)
    buf = C.create_string_buffer(128)
    st = L.b200_nccl_unique_id(buf)
    if st != abi.OK:
        raise B200Error(st, L.b200_last_error().decode())
    return buf.raw


def factor_eval(ftype, params, consts, xcat, whiten=False, device=0, so_path=None):
    """one factor instance on the device (b200_debug_factor_eval): (r[rows], J[rows][cols])"""
    L = load(so_path)
    rows, cols = abi.FACTOR_INFO[ftype][0], sum(abi.TAN_DIM[v] for v in abi.FACTOR_INFO[ftype][1])
    p = np.zeros(abi.MAX_FACTOR_PARAMS); p[:len(params)] = params
    c = np.ascontiguousarray(consts if consts is not None and len(consts) else np.zeros(1), dtype=np.float64)
    x = np.ascontiguousarray(xcat, dtype=np.float64)
    r = np.zeros(rows); J = np.zeros((rows, cols))
    st = L.b200_debug_factor_eval(C.c_int32(device), C.c_int32(ftype), D(p), D(c), D(x), C.c_int32(1 if whiten else 0), D(r), D(J))
    if st != abi.OK:
        raise B200Error(st, L.b200_last_error().decode())
    return r, J


def preintegrate(acc, gyro, dt, n_per, bias_hat, noise=None, combined=True, device=0, so_path=None):
    """Device preintegration (A0). acc, gyro: [n_imu][n_samples][3] -> records [n_imu][K-1][190|115]."""
    L = load(so_path)
    acc = np.ascontiguousarray(acc, dtype=np.float64); gyro = np.ascontiguousarray(gyro, dtype=np.float64)
    n_imu, n_samples, _ = acc.shape
    K = (n_samples + n_per - 1) // n_per
    n_int = K - 1
    rec = 190 if combined else 115
    out = np.zeros((n_imu, n_int, rec))
    noise = noise or abi.ImuNoise.opal_v2()
    bh = np.ascontiguousarray(bias_hat, dtype=np.float64)
    st = L.b200_preintegrate(C.c_int32(device), C.c_int32(n_imu), C.c_int32(n_samples), C.c_int32(n_per), C.c_int32(n_int), D(acc), D(gyro),
                             C.c_double(dt), D(bh), C.byref(noise), C.c_int32(1 if combined else 0), D(out))
    if st != abi.OK:
        raise B200Error(st, L.b200_last_error().decode())
    return out

"""ctypes binding of libppx_host.so -- the C++ host layer that mirrors the reference's driver surface
(alsCP_DT / alsCP_PP / CPD<>::als / hosvd / alsTucker_*; host/capi.cxx).  Used by tests/ and bench.py only.
Arrays cross as numpy float64 in the reference's global order (`x.ravel(order="F")`)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import HOST_LIB_PATH, PpxError, load_library

_vp, _i64 = C.c_void_p, C.c_int64
_hlib = None


def load_host_library():
    global _hlib
    if _hlib is not None:
        return _hlib
    load_library()  # libppx.so first (RTLD_GLOBAL)
    if not os.path.exists(HOST_LIB_PATH):
        raise PpxError(f"{HOST_LIB_PATH} is missing: build it with `make -C {os.path.dirname(HOST_LIB_PATH)}`")
    lib = C.CDLL(HOST_LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.ppxh_last_error.restype = C.c_char_p
    for name in ("ppxh_world_create", "ppxh_world_ctx", "ppxh_tensor_create", "ppxh_matrix_create", "ppxh_tensor_data",
                 "ppxh_cpd_create", "ppxh_cpd_create_lr", "ppxh_tucker_create", "ppxh_cp_pp_ops_build"):
        getattr(lib, name).restype = _vp
    lib.ppxh_world_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t]
    lib.ppxh_world_destroy.argtypes = [_vp]
    lib.ppxh_world_ctx.argtypes = [_vp]
    lib.ppxh_world_set.argtypes = [_vp, C.c_int, C.c_int]
    lib.ppxh_world_trim.argtypes = [_vp]
    lib.ppxh_world_comm_init.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, _i64, _i64, _i64]
    lib.ppxh_tensor_create.argtypes = [_vp, C.c_int, C.POINTER(_i64)]
    lib.ppxh_matrix_create.argtypes = [_vp, _i64, _i64]
    lib.ppxh_tensor_destroy.argtypes = [_vp]
    lib.ppxh_tensor_write.argtypes = [_vp, _vp]
    lib.ppxh_tensor_read.argtypes = [_vp, _vp]
    lib.ppxh_tensor_write_async.argtypes = [_vp, _vp]
    lib.ppxh_tensor_read_async.argtypes = [_vp, _vp]
    lib.ppxh_world_sync.argtypes = [_vp]
    lib.ppxh_tensor_data.argtypes = [_vp]
    lib.ppxh_tensor_size.argtypes = [_vp]
    lib.ppxh_tensor_size.restype = _i64
    lib.ppxh_tensor_order.argtypes = [_vp]
    lib.ppxh_tensor_len.argtypes = [_vp, C.c_int]
    lib.ppxh_tensor_len.restype = _i64
    lib.ppxh_tensor_norm2.argtypes = [_vp]
    lib.ppxh_tensor_norm2.restype = C.c_double
    lib.ppxh_tensor_fill.argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_double, C.c_double, _i64]
    lib.ppxh_build_V.argtypes = [_vp, C.POINTER(_vp), C.c_int, _vp]
    lib.ppxh_cp_residual.argtypes = [_vp, C.POINTER(_vp), C.c_int, _vp]
    lib.ppxh_cp_residual.restype = C.c_double
    lib.ppxh_trace_begin.argtypes = [C.c_int, C.c_int]
    lib.ppxh_trace_rows.argtypes = [_vp, C.c_int]
    lib.ppxh_trace_events.argtypes = [_vp, C.c_int]
    lib.ppxh_trace_sweeps.argtypes = [_vp, C.c_int]
    lib.ppxh_trace_bench.argtypes = [_vp, C.c_int]
    PV = C.POINTER(_vp)
    PI = C.POINTER(C.c_int)
    d = C.c_double
    lib.ppxh_alsCP.argtypes = [_vp, PV, PV, PV, C.c_int, d, d, C.c_int, _vp, PI]
    lib.ppxh_alsCP_DT.argtypes = [_vp, PV, PV, PV, C.c_int, d, d, C.c_int, d, C.c_char_p, C.c_int, C.c_int, _vp, PI]
    lib.ppxh_alsCP_PP.argtypes = [_vp, PV, PV, PV, C.c_int, d, d, d, C.c_int, d, d, C.c_char_p, C.c_int, C.c_int, _vp,
                                  PI]
    lib.ppxh_alsCP_PP_partupdate.argtypes = [_vp, PV, PV, PV, C.c_int, d, d, d, C.c_int, d, d, d, C.c_char_p, C.c_int,
                                             C.c_int, _vp, PI]
    lib.ppxh_cp_dt_sweeps.argtypes = [_vp, PV, PV, C.c_int, C.c_int, d, _vp]
    lib.ppxh_cp_pp_phase_timed.argtypes = [_vp, PV, PV, C.c_int, C.c_int, d, d, _vp, C.POINTER(C.c_float),
                                           C.POINTER(C.c_float)]
    lib.ppxh_cpd_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _vp]
    lib.ppxh_cpd_create_lr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
    lib.ppxh_cpd_destroy.argtypes = [_vp]
    lib.ppxh_cpd_init.argtypes = [_vp, _vp, PV, C.c_int, d, C.c_uint64]
    lib.ppxh_cpd_step.argtypes = [_vp, C.POINTER(d)]
    lib.ppxh_cpd_als.argtypes = [_vp, d, d, C.c_int, C.c_int, C.c_char_p, C.c_int, PI]
    lib.ppxh_cpd_read_W.argtypes = [_vp, C.c_int, _vp]
    lib.ppxh_cpd_read_grad.argtypes = [_vp, C.c_int, _vp]
    lib.ppxh_world_set_fast_residual.argtypes = [_vp, C.c_int]
    lib.ppxh_hosvd.argtypes = [_vp, _vp, PV, C.c_int, PI, _vp]
    lib.ppxh_alsTucker_DT.argtypes = [_vp, _vp, PV, C.c_int, d, d, C.c_int, C.c_char_p, C.c_int, C.c_int, _vp, PI]
    lib.ppxh_alsTucker_PP.argtypes = [_vp, _vp, PV, C.c_int, d, d, d, C.c_int, C.c_char_p, C.c_int, C.c_int, _vp, PI]
    lib.ppxh_alsTucker.argtypes = [_vp, _vp, PV, C.c_int, d, d, C.c_int, _vp, PI]
    lib.ppxh_cp_dt_mttkrps.argtypes = [_vp, PV, PV, C.c_int, _vp]
    lib.ppxh_cp_pp_ops_build.argtypes = [_vp, PV, C.c_int, _vp]
    lib.ppxh_cp_pp_ops_size.argtypes = [_vp, C.c_char_p]
    lib.ppxh_cp_pp_ops_size.restype = _i64
    lib.ppxh_cp_pp_ops_read.argtypes = [_vp, C.c_char_p, _vp]
    lib.ppxh_cp_pp_ops_free.argtypes = [_vp]
    lib.ppxh_tucker_create.argtypes = [C.c_int, PI, PI, _vp]
    lib.ppxh_tucker_destroy.argtypes = [_vp]
    lib.ppxh_tucker_init.argtypes = [_vp, _vp]
    lib.ppxh_tucker_als.argtypes = [_vp, C.c_int, d, d, d, C.c_int, C.c_int, C.c_char_p, PI]
    lib.ppxh_tucker_read_W.argtypes = [_vp, C.c_int, _vp]
    lib.ppxh_tucker_read_core.argtypes = [_vp, _vp]
    _hlib = lib
    return lib


def _ck(rc):
    if rc:
        raise PpxError(f"host layer error: {load_host_library().ppxh_last_error().decode()}")


class World:
    """CTF::World stand-in: one GPU, one stream.  solver: 0 Cholesky (default), 1 SVD pseudo-inverse semantics."""

    def __init__(self, device=0, solver=0, use_graph=True, workspace_bytes=1 << 30):
        self.lib = load_host_library()
        self.h = self.lib.ppxh_world_create(device, solver, int(use_graph), workspace_bytes)
        if not self.h:
            raise PpxError(f"cannot create World: {self.lib.ppxh_last_error().decode()} (no CPU fallback)")
        self.rank, self.np = 0, 1

    def set(self, solver=0, use_graph=True):
        self.lib.ppxh_world_set(self.h, solver, int(use_graph))

    def set_fast_residual(self, on):
        """alsCP_DT print points: residual from ||V||^2 - 2<M_N,W_N> + <S_N,G_N> instead of a pass over V."""
        self.lib.ppxh_world_set_fast_residual(self.h, int(on))

    def comm_init(self, id_bytes, nranks, rank, shard_mode, shard_global, row_begin, row_end):
        buf = C.create_string_buffer(bytes(id_bytes), 128) if id_bytes is not None else None
        _ck(self.lib.ppxh_world_comm_init(self.h, buf, nranks, rank, shard_mode, shard_global, row_begin, row_end))
        self.rank, self.np = rank, nranks

    def set_shard(self, shard_mode, shard_global, row_begin, row_end):
        self.lib.ppxh_world_set_shard.argtypes = [_vp, C.c_int, _i64, _i64, _i64]
        self.lib.ppxh_world_set_shard(self.h, shard_mode, shard_global, row_begin, row_end)

    def ctx_handle(self):
        return self.lib.ppxh_world_ctx(self.h)

    def launch_count(self):
        return int(load_library().ppx_launch_count(self.ctx_handle()))

    def sync(self):
        load_library().ppx_sync(self.ctx_handle())

    def trim(self):
        self.lib.ppxh_world_trim(self.h)

    def close(self):
        if self.h:
            self.lib.ppxh_world_destroy(self.h)
            self.h = None


class Tensor:
    def __init__(self, world, lens, matrix=False):
        self.world, self.lens, self.lib = world, tuple(int(v) for v in lens), world.lib
        if matrix:
            self.h = self.lib.ppxh_matrix_create(world.h, self.lens[0], self.lens[1])
        else:
            self.h = self.lib.ppxh_tensor_create(world.h, len(self.lens), (_i64 * len(self.lens))(*self.lens))
        if not self.h:
            raise PpxError(f"tensor allocation failed: {self.lib.ppxh_last_error().decode()}")

    @classmethod
    def from_numpy(cls, world, arr, matrix=False):
        t = cls(world, arr.shape, matrix)
        t.write(arr)
        return t

    def write(self, arr):
        flat = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).ravel(order="F"))
        assert flat.size == self.size()
        _ck(self.lib.ppxh_tensor_write(self.h, flat.ctypes.data_as(_vp)))

    def numpy(self):
        lens = tuple(self.lib.ppxh_tensor_len(self.h, i) for i in range(self.lib.ppxh_tensor_order(self.h)))
        out = np.empty(int(np.prod(lens)) if lens else 1, dtype=np.float64)
        _ck(self.lib.ppxh_tensor_read(self.h, out.ctypes.data_as(_vp)))
        return out.reshape(lens, order="F")

    def size(self):
        return int(self.lib.ppxh_tensor_size(self.h))

    def fill(self, seed, tensor_id, lo=0.0, hi=1.0, start=0):
        _ck(self.lib.ppxh_tensor_fill(self.h, seed, tensor_id, lo, hi, start))

    def norm2(self):
        return float(self.lib.ppxh_tensor_norm2(self.h))

    def data_ptr(self):
        return self.lib.ppxh_tensor_data(self.h)

    def free(self):
        if self.h:
            self.lib.ppxh_tensor_destroy(self.h)
            self.h = None


def Matrix(world, nrow, ncol):
    return Tensor(world, (nrow, ncol), matrix=True)


def _harr(ts):
    if ts is None:
        return None
    return (_vp * len(ts))(*[t.h for t in ts])


class Trace:
    """Rows the reference prints ([iter], [gradnorm], [pp_update], [diffV], [dtime]) and the switching markers."""

    def __init__(self, quiet=True, skip_residual=False):
        self.lib = load_host_library()
        self.quiet, self.skip = quiet, skip_residual

    def __enter__(self):
        self.lib.ppxh_trace_begin(int(self.quiet), int(self.skip))
        return self

    def __exit__(self, *a):
        n = self.lib.ppxh_trace_rows(None, 0)
        rows = np.zeros((max(n, 1), 5))
        self.lib.ppxh_trace_rows(rows.ctypes.data_as(_vp), n)
        self.rows = [tuple(r) for r in rows[:n]]
        for name in ("events", "sweeps"):
            fn = getattr(self.lib, "ppxh_trace_" + name)
            n = fn(None, 0)
            buf = np.zeros((max(n, 1), 2), dtype=np.int32)
            fn(buf.ctypes.data_as(_vp), n)
            setattr(self, name, [(int(a), int(b)) for a, b in buf[:n]])
        n = self.lib.ppxh_trace_bench(None, 0)
        b = np.zeros(max(n, 1))
        self.lib.ppxh_trace_bench(b.ctypes.data_as(_vp), n)
        self.bench_times = [float(x) for x in b[:n]]
        self.lib.ppxh_trace_end()
        return False


def build_V(world, V, W):
    _ck(world.lib.ppxh_build_V(V.h, _harr(W), len(W), world.h))


def cp_residual(world, V, W):
    return float(world.lib.ppxh_cp_residual(V.h, _harr(W), len(W), world.h))


def alsCP(world, V, W, grad_W, F, tol, maxiter, timelimit=5e3):
    st = C.c_int(0)
    _ck(world.lib.ppxh_alsCP(V.h, _harr(W), _harr(grad_W), _harr(F), len(W), tol, timelimit, maxiter, world.h,
                             C.byref(st)))
    return bool(st.value)


def alsCP_DT(world, V, W, grad_W, F, tol, maxiter, lam=0.0, resprint=10, bench=False, csv=None, timelimit=5e3):
    st = C.c_int(0)
    _ck(world.lib.ppxh_alsCP_DT(V.h, _harr(W), _harr(grad_W), _harr(F), len(W), tol, timelimit, maxiter, lam,
                                csv.encode() if csv else None, resprint, int(bench), world.h, C.byref(st)))
    return bool(st.value)


def alsCP_PP(world, V, W, grad_W, F, tol, tol_init, maxiter, lam=0.0, ratio_step=1.0, resprint=10, bench=False,
             csv=None, timelimit=5e3):
    st = C.c_int(0)
    _ck(world.lib.ppxh_alsCP_PP(V.h, _harr(W), _harr(grad_W), _harr(F), len(W), tol, tol_init, timelimit, maxiter, lam,
                                ratio_step, csv.encode() if csv else None, resprint, int(bench), world.h, C.byref(st)))
    return bool(st.value)


def alsCP_PP_partupdate(world, V, W, grad_W, F, tol, tol_init, maxiter, lam=0.0, ratio_step=1.0,
                        update_percentage=1.0, resprint=10, bench=False, csv=None, timelimit=5e3):
    st = C.c_int(0)
    _ck(world.lib.ppxh_alsCP_PP_partupdate(V.h, _harr(W), _harr(grad_W), _harr(F), len(W), tol, tol_init, timelimit,
                                           maxiter, lam, ratio_step, update_percentage, csv.encode() if csv else None,
                                           resprint, int(bench), world.h, C.byref(st)))
    return bool(st.value)


def cp_dt_sweeps(world, V, W, grad_W, n_sweeps, lam=0.0):
    """Exactly n exact ALS-DT sweeps on the world's stream, no logging, no host synchronisation (bench.py)."""
    _ck(world.lib.ppxh_cp_dt_sweeps(V.h, _harr(W), _harr(grad_W), len(W), n_sweeps, lam, world.h))


def cp_pp_phase_timed(world, V, W, grad_W, n_sweeps, lam=0.0, ratio_step=1.0):
    """One PP phase: operator build (ms) and n approximate sweeps (ms), CUDA events on the world's stream."""
    a, b = C.c_float(0), C.c_float(0)
    _ck(world.lib.ppxh_cp_pp_phase_timed(V.h, _harr(W), _harr(grad_W), len(W), n_sweeps, lam, ratio_step, world.h,
                                         C.byref(a), C.byref(b)))
    return a.value, b.value


def cp_dt_mttkrps(world, V, W):
    """MTTKRP of every mode from ONE pass over the dimension tree at fixed W (the calls alsCP_DT's sweep makes, without the
    updates in between); returns device matrices."""
    M = [Matrix(world, w.lens[0], w.lens[1]) for w in W]
    _ck(world.lib.ppxh_cp_dt_mttkrps(V.h, _harr(W), _harr(M), len(W), world.h))
    return M


class PPOperators:
    """The operators of one PP phase built at W (Build_mttkrp_map for all pairs, then all singles)."""

    def __init__(self, world, V, W):
        self.lib = world.lib
        self.h = self.lib.ppxh_cp_pp_ops_build(V.h, _harr(W), len(W), world.h)
        if not self.h:
            raise PpxError(self.lib.ppxh_last_error().decode())

    def get(self, key, shape):
        n = self.lib.ppxh_cp_pp_ops_size(self.h, key.encode())
        if n < 0:
            raise KeyError(key)
        out = np.empty(n)
        _ck(self.lib.ppxh_cp_pp_ops_read(self.h, key.encode(), out.ctypes.data_as(_vp)))
        return out.reshape(shape, order="F")

    def free(self):
        if self.h:
            self.lib.ppxh_cp_pp_ops_free(self.h)
            self.h = None


class CPD:
    """CPD<double, Optimizer> (src/CP.h); kind: 'simple' | 'dt' | 'msdt' | 'dtlr' | 'msdtlr' (the last two take
    update_rank and randomsvd, run.cxx -pp 2 / 3)."""

    KINDS = {"simple": 0, "dt": 1, "msdt": 2, "dtlr": 3, "msdtlr": 4}

    def __init__(self, world, kind, order, size, r, update_rank=1, randomsvd=0):
        self.world, self.lib, self.order = world, world.lib, order
        if self.KINDS[kind] >= 3:
            self.h = self.lib.ppxh_cpd_create_lr(self.KINDS[kind], order, size, r, update_rank, randomsvd, world.h)
        else:
            self.h = self.lib.ppxh_cpd_create(self.KINDS[kind], order, size, r, world.h)
        if not self.h:
            raise PpxError(self.lib.ppxh_last_error().decode())

    def Init(self, V, W, lam=0.0, grad_seed=3):
        self.shapes = [w.lens for w in W]
        _ck(self.lib.ppxh_cpd_init(self.h, V.h, _harr(W), len(W), lam, grad_seed))

    def step(self):
        f = C.c_double(0)
        _ck(self.lib.ppxh_cpd_step(self.h, C.byref(f)))
        return f.value

    def als(self, tol, maxsweep, resprint=10, bench=False, csv=None, timelimit=5e3):
        st = C.c_int(0)
        _ck(self.lib.ppxh_cpd_als(self.h, tol, timelimit, maxsweep, resprint, csv.encode() if csv else None, int(bench),
                                  C.byref(st)))
        return bool(st.value)

    def W(self, i):
        out = np.empty(self.shapes[i][0] * self.shapes[i][1])
        _ck(self.lib.ppxh_cpd_read_W(self.h, i, out.ctypes.data_as(_vp)))
        return out.reshape(self.shapes[i], order="F")

    def grad(self, i):
        out = np.empty(self.shapes[i][0] * self.shapes[i][1])
        _ck(self.lib.ppxh_cpd_read_grad(self.h, i, out.ctypes.data_as(_vp)))
        return out.reshape(self.shapes[i], order="F")

    def free(self):
        if self.h:
            self.lib.ppxh_cpd_destroy(self.h)
            self.h = None


def hosvd(world, V, core, W, ranks):
    _ck(world.lib.ppxh_hosvd(V.h, core.h, _harr(W), len(W), (C.c_int * len(ranks))(*ranks), world.h))


def alsTucker_DT(world, V, core, W, tol, maxiter, resprint=10, bench=False, csv=None, timelimit=5e3):
    st = C.c_int(0)
    _ck(world.lib.ppxh_alsTucker_DT(V.h, core.h, _harr(W), len(W), tol, timelimit, maxiter,
                                    csv.encode() if csv else None, resprint, int(bench), world.h, C.byref(st)))
    return bool(st.value)


def alsTucker_PP(world, V, core, W, tol, tol_init, maxiter, resprint=10, bench=False, csv=None, timelimit=5e3):
    st = C.c_int(0)
    _ck(world.lib.ppxh_alsTucker_PP(V.h, core.h, _harr(W), len(W), tol, tol_init, timelimit, maxiter,
                                    csv.encode() if csv else None, resprint, int(bench), world.h, C.byref(st)))
    return bool(st.value)


def alsTucker(world, V, core, W, tol, maxiter, timelimit=5e3):
    st = C.c_int(0)
    _ck(world.lib.ppxh_alsTucker(V.h, core.h, _harr(W), len(W), tol, timelimit, maxiter, world.h, C.byref(st)))
    return bool(st.value)


class Tucker:
    """Tucker<double> (src/Tucker.h): HOSVD initialisation at Init, then als (alsTucker_DT) or als_pp (alsTucker_PP)."""

    def __init__(self, world, sizes, ranks):
        self.world, self.lib = world, world.lib
        self.sizes, self.ranks = list(sizes), list(ranks)
        n = len(self.sizes)
        self.h = self.lib.ppxh_tucker_create(n, (C.c_int * n)(*self.sizes), (C.c_int * n)(*self.ranks), world.h)
        if not self.h:
            raise PpxError(self.lib.ppxh_last_error().decode())

    def Init(self, V):
        _ck(self.lib.ppxh_tucker_init(self.h, V.h))

    def als(self, tol, maxiter, resprint=10, pp=False, tol_init=1e-2, csv=None, timelimit=5e3):
        st = C.c_int(0)
        _ck(self.lib.ppxh_tucker_als(self.h, int(pp), tol, tol_init, timelimit, maxiter, resprint,
                                     csv.encode() if csv else None, C.byref(st)))
        return bool(st.value)

    def W(self, i):
        out = np.empty(self.sizes[i] * self.ranks[i])
        _ck(self.lib.ppxh_tucker_read_W(self.h, i, out.ctypes.data_as(_vp)))
        return out.reshape((self.sizes[i], self.ranks[i]), order="F")

    def core(self):
        out = np.empty(int(np.prod(self.ranks)))
        _ck(self.lib.ppxh_tucker_read_core(self.h, out.ctypes.data_as(_vp)))
        return out.reshape(tuple(self.ranks), order="F")

    def free(self):
        if self.h:
            self.lib.ppxh_tucker_destroy(self.h)
            self.h = None

"""pairwise-perturbation_b200 -- B200-native engine for the ALS / pairwise-perturbation sweeps of CP and Tucker.

The product is native: `libppx.so` (hand-written sm_100a CUDA behind the C ABI of include/ppx.h) and
`libppx_host.so` (the C++ host layer mirroring the reference's als_CP.h / als_Tucker.h / src/CP.h surface, plus the
`test_ALS` / `pp_bench` / `run` command lines).  This Python module is only the ctypes binding used by tests/ and
bench.py: it moves no data through numpy on the compute path and has NO CPU fallback -- if the CUDA library is
missing or no GPU is present every compute entry point raises.

The directory name contains a hyphen (it mirrors the reference's repository name), so import it with
`importlib.import_module("pairwise-perturbation_b200")`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPX_LIB_PATH") or os.path.join(_HERE, "libppx.so")  # override: A/B tests of kernel builds
HOST_LIB_PATH = os.path.join(_HERE, "libppx_host.so")

PPX_SOLVE_CHOL = 0
PPX_SOLVE_SVD_PINV = 1

_i64 = C.c_int64
_dp = C.c_void_p  # device pointer
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/ppx.h
SIGNATURES = {
    "ppx_version": (C.c_char_p, []),
    "ppx_ctx_create": (C.c_int, [C.c_int, _vp, C.c_size_t, C.POINTER(_vp)]),
    "ppx_ctx_destroy": (C.c_int, [_vp]),
    "ppx_sync": (C.c_int, [_vp]),
    "ppx_last_error": (C.c_char_p, [_vp]),
    "ppx_stream": (_vp, [_vp]),
    "ppx_device": (C.c_int, [_vp]),
    "ppx_sm_count": (C.c_int, [_vp]),
    "ppx_launch_count": (_i64, [_vp]),
    "ppx_malloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "ppx_free": (C.c_int, [_vp, _dp]),
    "ppx_host_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "ppx_host_free": (C.c_int, [_vp, _vp]),
    "ppx_memcpy_h2d": (C.c_int, [_vp, _dp, _vp, C.c_size_t]),
    "ppx_memcpy_d2h": (C.c_int, [_vp, _vp, _dp, C.c_size_t]),
    "ppx_memcpy_d2d": (C.c_int, [_vp, _dp, _dp, C.c_size_t]),
    "ppx_memcpy2d_d2d": (C.c_int, [_vp, _dp, C.c_size_t, _dp, C.c_size_t, C.c_size_t, C.c_size_t]),
    "ppx_memset_zero": (C.c_int, [_vp, _dp, C.c_size_t]),
    "ppx_mem_info": (C.c_int, [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "ppx_event_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "ppx_event_destroy": (C.c_int, [_vp, _vp]),
    "ppx_event_record": (C.c_int, [_vp, _vp]),
    "ppx_event_elapsed_ms": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_float)]),
    "ppx_graph_begin": (C.c_int, [_vp]),
    "ppx_graph_end": (C.c_int, [_vp, C.POINTER(_vp)]),
    "ppx_graph_launch": (C.c_int, [_vp, _vp]),
    "ppx_graph_destroy": (C.c_int, [_vp, _vp]),
    "ppx_fill_uniform": (C.c_int, [_vp, _dp, _i64, C.c_uint64, C.c_uint64, _i64, C.c_double, C.c_double]),
    "ppx_fill_uniform_rows": (C.c_int, [_vp, _dp, _i64, _i64, _i64, _i64, C.c_uint64, C.c_uint64, C.c_double,
                                        C.c_double]),
    "ppx_fill_laplacian": (C.c_int, [_vp, _dp, C.c_int, _i64]),
    "ppx_ttm_first": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, _dp, _i64, C.c_int, _dp]),
    "ppx_ttm_multi": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, C.c_int, C.POINTER(_dp), C.POINTER(_i64),
                                C.c_int, _dp]),
    "ppx_mttv": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, _dp, _i64, C.c_int, _dp]),
    "ppx_mttv3": (C.c_int, [_vp, _vp, C.POINTER(_i64), C.c_int, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "ppx_mttv2": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, _dp, _i64, C.c_int, _dp, _i64, C.c_int, _dp]),
    "ppx_ttm_first_mttv": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, _dp, _i64, C.c_int, _dp, _i64,
                                     C.c_int, _dp]),
    "ppx_pp_correct": (C.c_int, [_vp, _dp, C.POINTER(_dp), C.POINTER(C.c_int), C.POINTER(_dp), C.POINTER(_i64),
                                 C.c_int, _i64, C.c_int, _dp]),
    "ppx_gram": (C.c_int, [_vp, _dp, _i64, _i64, C.c_int, _dp]),
    "ppx_hadamard_grams": (C.c_int, [_vp, C.POINTER(_dp), C.c_int, C.c_int, C.c_int, C.c_double, _dp]),
    "ppx_solve_update": (C.c_int, [_vp, _dp, _dp, _dp, _i64, C.c_int, _dp, C.c_double, C.c_int, _dp, _dp, _dp]),
    "ppx_solve_update_g": (C.c_int, [_vp, _dp, C.POINTER(_dp), C.c_int, C.c_int, C.c_double, _dp, _i64, C.c_int, _dp,
                                     C.c_double, C.c_int, _dp, _dp, _dp]),
    "ppx_spd_inverse_g": (C.c_int, [_vp, C.POINTER(_dp), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _dp, _dp]),
    "ppx_solve_apply": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _i64, C.c_int, _dp, C.c_double, _dp, _dp]),
    "ppx_spd_factor_inverse": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "ppx_gemm_small": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, _i64, _dp, _i64,
                                 C.c_double, _dp, _i64]),
    "ppx_rank_expand_acc": (C.c_int, [_vp, _dp, _i64, C.c_int, _dp, _i64, C.c_int, _dp]),
    "ppx_side_begin": (C.c_int, [_vp]),
    "ppx_side_end": (C.c_int, [_vp]),
    "ppx_side_join": (C.c_int, [_vp]),
    "ppx_lane_begin": (C.c_int, [_vp, C.c_int]),
    "ppx_lane_end": (C.c_int, [_vp]),
    "ppx_lane_join": (C.c_int, [_vp, C.c_int]),
    "ppx_stamp": (C.c_int, [_vp, _vp]),
    "ppx_normalize_norms": (C.c_int, [_vp, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_i64), C.c_int, C.c_int,
                                      C.POINTER(_dp), _dp]),
    "ppx_normalize": (C.c_int, [_vp, C.POINTER(_dp), C.POINTER(_i64), C.c_int, C.c_int, C.POINTER(_dp)]),
    "ppx_normalize_g": (C.c_int, [_vp, C.POINTER(_dp), C.POINTER(_i64), C.c_int, C.c_int, C.POINTER(_dp)]),
    "ppx_sqnorms": (C.c_int, [_vp, C.POINTER(_dp), C.POINTER(_i64), C.c_int, _dp]),
    "ppx_dots": (C.c_int, [_vp, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_i64), C.c_int, _dp]),
    "ppx_diff_update": (C.c_int, [_vp, _dp, _dp, _dp, _i64, _dp]),
    "ppx_axpby": (C.c_int, [_vp, C.c_double, _dp, C.c_double, _dp, _i64]),
    "ppx_cp_residual": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.POINTER(_dp), C.c_int, _dp]),
    "ppx_cp_reconstruct": (C.c_int, [_vp, C.POINTER(_i64), C.c_int, C.POINTER(_dp), C.c_int, _dp]),
    "ppx_ttm": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, _dp, _i64, C.c_int, _dp]),
    "ppx_ttm_acc": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, _dp, _i64, C.c_int, _dp]),
    "ppx_unfold_gram": (C.c_int, [_vp, _dp, C.POINTER(_i64), C.c_int, C.c_int, _dp]),
    "ppx_sym_eig_topk": (C.c_int, [_vp, _dp, _i64, C.c_int, _dp, _dp]),
    "ppx_sym_eig_topk_warm": (C.c_int, [_vp, _dp, _i64, C.c_int, _dp, _dp, _dp, C.c_int]),
    "ppx_sign_align": (C.c_int, [_vp, _dp, _dp, _i64, C.c_int]),
    "ppx_diff_sqnorm": (C.c_int, [_vp, _dp, _dp, _i64, _dp]),
    "ppx_transpose": (C.c_int, [_vp, _dp, _i64, _i64, _dp]),
    "ppx_shard_range": (C.c_int, [_i64, C.c_int, C.c_int, C.POINTER(_i64), C.POINTER(_i64)]),
    "ppx_comm_unique_id": (C.c_int, [_vp]),
    "ppx_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "ppx_comm_bootstrap": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]),
    "ppx_comm_size": (C.c_int, [_vp]),
    "ppx_comm_rank": (C.c_int, [_vp]),
    "ppx_comm_p2p": (C.c_int, [_vp]),
    "ppx_allreduce_packed": (C.c_int, [_vp, C.POINTER(_dp), C.POINTER(_i64), C.c_int]),
    "ppx_alltoallv": (C.c_int, [_vp, C.POINTER(_dp), C.POINTER(_i64), C.POINTER(_dp), C.POINTER(_i64)]),
}

_lib = None


class PpxError(RuntimeError):
    pass


def load_library():
    """dlopen libppx.so (built in-tree by `make -C pairwise-perturbation_b200` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PPX_LIB", LIB_PATH)  # PPX_LIB: A/B timing of two builds of the library (tools/)
    if not os.path.exists(path):
        raise PpxError(f"{path} is missing: build it with `make -C {_HERE}`; there is no CPU fallback")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # raises AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def shard_range(s: int, nranks: int, rank: int):
    """Rows [begin, end) of a mode of size s owned by `rank` (pure host function of the C ABI)."""
    lib = load_library()
    b, e = _i64(), _i64()
    rc = lib.ppx_shard_range(s, nranks, rank, C.byref(b), C.byref(e))
    if rc:
        raise PpxError(f"ppx_shard_range({s},{nranks},{rank}) -> {rc}")
    return b.value, e.value


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _lens(lens):
    return (_i64 * len(lens))(*[int(v) for v in lens])


def _ptrs(ts):
    return (_dp * len(ts))(*[(t.data_ptr() if t is not None else None) for t in ts])


class Ctx:
    """One context per GPU (replaces CTF::World).  Device buffers are torch float64 CUDA tensors holding the raw
    first-index-fastest data; torch is used for memory only."""

    def __init__(self, device: int = 0, workspace_bytes: int = 256 << 20, stream=None):
        import torch

        self.lib = load_library()
        if not torch.cuda.is_available():
            raise PpxError("no CUDA device: pairwise-perturbation_b200 has no CPU fallback")
        torch.cuda.set_device(device)
        self.torch = torch
        self.device = torch.device("cuda", device)
        if stream is None:
            self._tstream = torch.cuda.Stream(device=self.device)
        else:
            self._tstream = stream
        h = _vp()
        rc = self.lib.ppx_ctx_create(device, C.c_void_p(self._tstream.cuda_stream), workspace_bytes, C.byref(h))
        if rc:
            raise PpxError(f"ppx_ctx_create failed ({rc})")
        self.h = h

    # -- plumbing -----------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.ppx_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise PpxError(f"ppx error {rc}: {self.lib.ppx_last_error(self.h).decode()}")

    def sync(self):
        self._ck(self.lib.ppx_sync(self.h))

    @property
    def stream(self):
        return self._tstream

    def launch_count(self):
        return int(self.lib.ppx_launch_count(self.h))

    def empty(self, n):
        with self.torch.cuda.stream(self._tstream):
            return self.torch.empty(int(n), dtype=self.torch.float64, device=self.device)

    def zeros(self, n):
        with self.torch.cuda.stream(self._tstream):
            return self.torch.zeros(int(n), dtype=self.torch.float64, device=self.device)

    def to_device(self, arr):
        """numpy array (indexed [a,b,..]) -> flat device buffer in first-index-fastest order."""
        import numpy as np

        flat = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).ravel(order="F"))
        with self.torch.cuda.stream(self._tstream):
            return self.torch.from_numpy(flat).to(self.device)

    def to_host(self, t, shape):
        self.sync()
        return t.detach().cpu().numpy().reshape(shape, order="F")

    # -- operators ----------------------------------------------------------------------------------------------
    def fill_uniform(self, out, seed, tensor_id, start=0, lo=0.0, hi=1.0):
        self._ck(self.lib.ppx_fill_uniform(self.h, _ptr(out), out.numel(), seed, tensor_id, start, lo, hi))

    def ttm_first(self, V, lens, x, W, R, out, ldw=None):
        self._ck(self.lib.ppx_ttm_first(self.h, _ptr(V), _lens(lens), len(lens), x, _ptr(W),
                                        ldw if ldw is not None else lens[x], R, _ptr(out)))

    def ttm_multi(self, V, lens, x_first, Ws, R, out):
        n = len(Ws)
        self._ck(self.lib.ppx_ttm_multi(self.h, _ptr(V), _lens(lens), len(lens), x_first, n, _ptrs(Ws),
                                        _lens([lens[x_first + j] for j in range(n)]), R, _ptr(out)))

    def mttv(self, T, lens, x, W, R, out, ldw=None):
        self._ck(self.lib.ppx_mttv(self.h, _ptr(T), _lens(lens), len(lens), x, _ptr(W),
                                   ldw if ldw is not None else lens[x], R, _ptr(out)))

    def mttv3(self, T, lens3, Wl, Wx, Wt, R, out_l, out_x, out_t):
        """All (non-None) Hadamard contractions of one order-3 (+rank) tensor in one pass (ppx_mttv3)."""
        def p(x):
            return _ptr(x) if x is not None else None
        self._ck(self.lib.ppx_mttv3(self.h, _ptr(T), _lens(lens3), R, p(Wl), lens3[0], p(Wx), lens3[1], p(Wt), lens3[2],
                                    p(out_l), p(out_x), p(out_t)))

    def mttv2(self, T, lens, x1, W1, x2, W2, R, out):
        self._ck(self.lib.ppx_mttv2(self.h, _ptr(T), _lens(lens), len(lens), x1, _ptr(W1), lens[x1], x2, _ptr(W2),
                                    lens[x2], R, _ptr(out)))

    def ttm_first_mttv(self, V, lens, x1, W1, x2, W2, R, out):
        self._ck(self.lib.ppx_ttm_first_mttv(self.h, _ptr(V), _lens(lens), len(lens), x1, _ptr(W1), lens[x1], x2,
                                             _ptr(W2), lens[x2], R, _ptr(out)))

    def pp_correct(self, M0, ops, which, dWs, s_other, s_i, R, out):
        n = len(ops)
        self._ck(self.lib.ppx_pp_correct(self.h, _ptr(M0), _ptrs(ops), (C.c_int * n)(*which), _ptrs(dWs),
                                         _lens(s_other), n, s_i, R, _ptr(out)))

    def gram(self, W, s, R, G, ldw=None):
        self._ck(self.lib.ppx_gram(self.h, _ptr(W), s, ldw if ldw is not None else s, R, _ptr(G)))

    def spd_inverse_g(self, Gs, skip, lam, R, mode, S_out, Sinv_out):
        self._ck(self.lib.ppx_spd_inverse_g(self.h, _ptrs(Gs), len(Gs), skip, lam, R, mode,
                                            _ptr(S_out) if S_out is not None else None, _ptr(Sinv_out)))

    def hadamard_grams(self, Gs, skip, R, lam, S):
        self._ck(self.lib.ppx_hadamard_grams(self.h, _ptrs(Gs), len(Gs), skip, R, lam, _ptr(S)))

    def solve_update(self, M, S, W, s, R, W_init=None, ratio_step=1.0, mode=PPX_SOLVE_CHOL, grad=None, dW=None,
                     sq_norms=None):
        self._ck(self.lib.ppx_solve_update(self.h, _ptr(M), _ptr(S), _ptr(W), s, R, _ptr(W_init), ratio_step, mode,
                                           _ptr(grad), _ptr(dW), _ptr(sq_norms)))

    def solve_update_g(self, M, Gs, skip, lam, W, s, R, W_init=None, ratio_step=1.0, mode=PPX_SOLVE_CHOL, grad=None,
                       dW=None, sq_norms=None):
        self._ck(self.lib.ppx_solve_update_g(self.h, _ptr(M), _ptrs(Gs), len(Gs), skip, lam, _ptr(W), s, R,
                                             _ptr(W_init), ratio_step, mode, _ptr(grad), _ptr(dW), _ptr(sq_norms)))

    def normalize(self, Ws, sizes, R, Gs=None):
        self._ck(self.lib.ppx_normalize(self.h, _ptrs(Ws), _lens(sizes), len(Ws), R,
                                        _ptrs(Gs) if Gs is not None else None))

    def normalize_g(self, Ws, sizes, R, Gs):
        self._ck(self.lib.ppx_normalize_g(self.h, _ptrs(Ws), _lens(sizes), len(Ws), R, _ptrs(Gs)))

    def normalize_norms(self, Ws, dWs, sizes, R, Gs, sq_out):
        self._ck(self.lib.ppx_normalize_norms(self.h, _ptrs(Ws), _ptrs(dWs) if dWs is not None else None, _lens(sizes),
                                              len(Ws), R, _ptrs(Gs), _ptr(sq_out)))

    def sqnorms(self, Xs, out):
        self._ck(self.lib.ppx_sqnorms(self.h, _ptrs(Xs), _lens([x.numel() for x in Xs]), len(Xs), _ptr(out)))

    def diff_update(self, W, W_prev, dW, sq_out):
        self._ck(self.lib.ppx_diff_update(self.h, _ptr(W), _ptr(W_prev), _ptr(dW), W.numel(), _ptr(sq_out)))

    def axpby(self, alpha, x, beta, y):
        self._ck(self.lib.ppx_axpby(self.h, alpha, _ptr(x), beta, _ptr(y), y.numel()))

    def cp_residual(self, V, lens, Ws, R, sq_out):
        self._ck(self.lib.ppx_cp_residual(self.h, _ptr(V), _lens(lens), len(lens), _ptrs(Ws), R, _ptr(sq_out)))

    def cp_reconstruct(self, lens, Ws, R, V_out):
        self._ck(self.lib.ppx_cp_reconstruct(self.h, _lens(lens), len(lens), _ptrs(Ws), R, _ptr(V_out)))

    def spd_factor_inverse(self, S, R, Linv):
        self._ck(self.lib.ppx_spd_factor_inverse(self.h, _ptr(S), R, _ptr(Linv)))

    def gemm_small(self, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
        self._ck(self.lib.ppx_gemm_small(self.h, int(ta), int(tb), m, n, k, alpha, _ptr(A), lda, _ptr(B), ldb, beta,
                                         _ptr(Cm), ldc))

    def rank_expand_acc(self, T, Mtot, r, VT, ldvt, R, out):
        self._ck(self.lib.ppx_rank_expand_acc(self.h, _ptr(T), Mtot, r, _ptr(VT), ldvt, R, _ptr(out)))

    def fill_laplacian(self, out, d, s):
        self._ck(self.lib.ppx_fill_laplacian(self.h, _ptr(out), d, s))

    def ttm(self, T, lens, x, W, Q, out, acc=False, ldw=None):
        fn = self.lib.ppx_ttm_acc if acc else self.lib.ppx_ttm
        self._ck(fn(self.h, _ptr(T), _lens(lens), len(lens), x, _ptr(W), ldw if ldw is not None else lens[x], Q,
                    _ptr(out)))

    def unfold_gram(self, T, lens, i, MTM):
        self._ck(self.lib.ppx_unfold_gram(self.h, _ptr(T), _lens(lens), len(lens), i, _ptr(MTM)))

    def sym_eig_topk(self, MTM, s, r, U, evals=None, basis=None, basis_valid=False):
        if basis is None:
            self._ck(self.lib.ppx_sym_eig_topk(self.h, _ptr(MTM), s, r, _ptr(U), _ptr(evals)))
        else:
            self._ck(self.lib.ppx_sym_eig_topk_warm(self.h, _ptr(MTM), s, r, _ptr(U), _ptr(evals), _ptr(basis),
                                                    1 if basis_valid else 0))

    def sign_align(self, U, Uref, s, r):
        self._ck(self.lib.ppx_sign_align(self.h, _ptr(U), _ptr(Uref), s, r))

    def diff_sqnorm(self, a, b, sq_out):
        self._ck(self.lib.ppx_diff_sqnorm(self.h, _ptr(a), _ptr(b), a.numel(), _ptr(sq_out)))

    def transpose(self, A, m, n, B):
        self._ck(self.lib.ppx_transpose(self.h, _ptr(A), m, n, _ptr(B)))

    # -- collectives --------------------------------------------------------------------------------------------
    def comm_init(self, id_bytes, nranks, rank):
        buf = C.create_string_buffer(bytes(id_bytes), 128) if id_bytes is not None else None
        self._ck(self.lib.ppx_comm_init(self.h, buf, nranks, rank))

    def allreduce_packed(self, bufs):
        self._ck(self.lib.ppx_allreduce_packed(self.h, _ptrs(bufs), _lens([b.numel() for b in bufs]), len(bufs)))


def comm_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.ppx_comm_unique_id(buf)
    if rc:
        raise PpxError(f"ppx_comm_unique_id failed ({rc})")
    return buf.raw

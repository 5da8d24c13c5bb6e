"""ctypes binding of libstcat_sm100.so (include/stcat_b200.h) and the tensor-level backend object.

There is deliberately NO fallback here: if the library is missing, or a tensor is not on a CUDA
device, the call raises.  ``tests/emu_backend.py`` provides a torch-CPU emulation of the same
methods for *testing the host-side composition without a GPU*; it is installed only by tests via
``ops.set_backend`` and never selected by product code.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from typing import Optional

import torch

LIB_NAME = "libstcat_sm100.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", LIB_NAME)

F32, BF16 = 0, 1

#: every symbol declared in include/stcat_b200.h: name -> (restype, argtypes)
_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float
SIGNATURES = {
    "stcat_abi_version": (c_int, []),
    "stcat_last_error": (c_char_p, []),
    "stcat_device_arch": (c_int, []),
    "stcat_linear_fwd": (c_int, [_P, _L, _I, _P, _L, _I, _P, _P, _L, _I, _I, _I, _I, _I, _I, _P]),
    "stcat_linear_bwd_data": (c_int, [_P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _I, _I, _I, _I, _P]),
    "stcat_linear_bwd_weight": (c_int, [_P, _L, _I, _P, _L, _I, _P, _L, _P, _I, _I, _I, _I, _P]),
    "stcat_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "stcat_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "stcat_attention_fwd": (c_int, [_P, _P, _L, _P, _P, _L, _P, _L, _P, _L, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "stcat_attention_bwd": (c_int, [_P, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _L, _I, _P, _P, _P, _P, _P, _P, _L, _P, _P,
                                    _L, _P, _L, _I, _I, _I, _I, _I, _F, _P]),
    "stcat_add": (c_int, [_P, _P, _P, _P, _L, _P]),
    "stcat_relu_bwd": (c_int, [_P, _I, _P, _I, _L, _P]),
    "stcat_cast_bf16": (c_int, [_P, _P, _L, _L, _I, _P]),
    "stcat_sted_score": (c_int, [_P, _P, _P, _P, _I, _I, _P]),
    "stcat_map2d_pool": (c_int, [_P, _P, _P, _I, _I, _I, _P]),
}



class _Term(ctypes.Structure):
    _fields_ = [("a", c_void_p), ("lda", c_int64), ("b", c_void_p), ("ldb", c_int64), ("bias", c_void_p),
                ("k", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class _Job(ctypes.Structure):
    _fields_ = [("term", _Term * 3), ("nterms", ctypes.c_int32), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
                ("relu", ctypes.c_int32), ("accumulate", ctypes.c_int32), ("out_dtype", ctypes.c_int32),
                ("out", c_void_p), ("ldo", c_int64), ("dbias", c_void_p)]


_U64 = ctypes.c_uint64
SIGNATURES["stcat_dropout"] = (c_int, [_P, _P, _I, _L, _F, _U64, _U64, _P])
SIGNATURES["stcat_linear_dropout_fwd"] = (c_int, [_P, _L, _I, _P, _L, _I, _P, _P, _L, _I, _I, _I, _I, _I, _F, _U64, _U64, _P])
SIGNATURES["stcat_linear_bwd_data_scaled"] = (c_int, [_P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _I, _P, _I, _I, _I, _F, _P])
SIGNATURES["stcat_layernorm_dropout_fwd"] = (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _U64, _U64, _P])
SIGNATURES["stcat_layernorm_dropout_bwd"] = (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _U64, _U64, _P])
SIGNATURES["stcat_attention_dropout_fwd"] = (c_int, [_P, _P, _L, _P, _P, _L, _P, _L, _P, _L, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F,
                                                     _F, _U64, _U64, _P, _I, _P])
SIGNATURES["stcat_attention_dropout_bwd"] = (c_int, [_P, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _L, _I, _P, _P, _P, _P, _P, _P, _L,
                                                     _P, _P, _L, _P, _L, _I, _I, _I, _I, _I, _F, _F, _U64, _U64, _P, _I, _P])
SIGNATURES["stcat_dropout_bits"] = (c_int, [_P, _L, _I, _I, _F, _U64, _U64, _P])
SIGNATURES["stcat_sumsq"] = (c_int, [_P, _L, _P, _P])
SIGNATURES["stcat_adamw_step"] = (c_int, [_P, _P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _L, _P, _F, _F, _P])
SIGNATURES["stcat_debug_attn_trace"] = (c_int, [_P])
SIGNATURES["stcat_debug_gemm_trace"] = (c_int, [_P])
SIGNATURES["stcat_set_dropout_step"] = (c_int, [_P])
SIGNATURES["stcat_set_gemm_sm_limit"] = (c_int, [_I])
SIGNATURES["stcat_set_sm_cap"] = (c_int, [_I])
SIGNATURES["stcat_pos_sine"] = (c_int, [_P, _P, _I, _I, _I, _I, _F, _F, _P])
SIGNATURES["stcat_box_interp"] = (c_int, [_P, _P, _I, _P, _L, _I, _P])
SIGNATURES["stcat_debug_attn_counts"] = (c_int, [ctypes.POINTER(ctypes.c_longlong)])
SIGNATURES["stcat_anchor_sine_fwd"] = (c_int, [_P, _P, _P, _L, _P])
SIGNATURES["stcat_anchor_sine_bwd"] = (c_int, [_P, _P, _P, _L, _P])
SIGNATURES["stcat_box_refine_fwd"] = (c_int, [_P, _P, _P, _L, _F, _P])
SIGNATURES["stcat_box_refine_bwd"] = (c_int, [_P, _P, _P, _P, _P, _L, _F, _P])
SIGNATURES["stcat_stg_loss"] = (c_int, [_P] * 12 + [ctypes.POINTER(c_float), _F, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P])
SIGNATURES["stcat_linear_group"] = (c_int, [_I, _I, ctypes.POINTER(_Job), _I, _P])
SIGNATURES["stcat_token_assembly"] = (c_int, [_P] * 10 + [_I] * 5 + [_P])
SIGNATURES["stcat_token_assembly_bwd"] = (c_int, [_P] * 5 + [_I] * 5 + [_P])
SIGNATURES["stcat_mem_operands"] = (c_int, [_P] * 6 + [_I] * 3 + [_P])
SIGNATURES["stcat_mem_operands_bwd"] = (c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P])
SIGNATURES["stcat_template_fwd"] = (c_int, [_P] * 17 + [_I] * 4 + [_P])
SIGNATURES["stcat_template_bwd"] = (c_int, [_P] * 27 + [_I] * 4 + [_P])
SIGNATURES["stcat_box_head_fwd"] = (c_int, [_P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P])
SIGNATURES["stcat_box_head_bwd"] = (c_int, [_P, _P, _P, _P, _P, _L, _P, _P, _P, _I, _I, _F, _P])
SIGNATURES["stcat_mul_cast"] = (c_int, [_P, _L, _P, _P, _P, _P, _P, _L, _I, _P])
SIGNATURES["stcat_mul_cast_bwd"] = (c_int, [_P, _I, _P, _L, _P, _L, _I, _P])
SIGNATURES["stcat_cls_gather"] = (c_int, [_P] * 6 + [_I] * 4 + [_P])
SIGNATURES["stcat_cls_scatter"] = (c_int, [_P] * 5 + [_I] * 4 + [_P])
MAX_GROUP_JOBS = 12
ABI_VERSION = 11  # include/stcat_b200.h STCAT_ABI_VERSION

_lib = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen the kernel library and bind every symbol of the header.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("STCAT_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{LIB_NAME} not found at {path}: build it with `python -m stcat_b200.build` "
            "(there is no CPU / PyTorch fallback for the STCAT hot path)")
    lib = ctypes.CDLL(path)
    lib.stcat_abi_version.restype = c_int
    got = lib.stcat_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"{path} implements C-ABI version {got}, this package expects {ABI_VERSION}: rebuild it with "
                           "`python -m stcat_b200.build --force`")
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class StcatError(RuntimeError):
    pass


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise StcatError(f"unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class CudaBackend:
    """Tensor-level calls into the C ABI on the current CUDA stream."""

    name = "cuda"

    def __init__(self):
        self.lib = load_library()
        self.launches = 0  # kernels launched through this backend (bench.py reports it)

    # -- helpers -------------------------------------------------------
    def _rc(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.stcat_last_error()
            raise StcatError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def set_dropout_step(self, counter: Optional[torch.Tensor]):
        """counter: one int64 on the device (kept alive by the caller) or None"""
        if counter is not None and (not counter.is_cuda or counter.dtype != torch.int64 or counter.numel() != 1):
            raise StcatError("dropout step counter: one int64 CUDA element")
        self._rc(self.lib.stcat_set_dropout_step(None if counter is None else counter.data_ptr()), "set_dropout_step")

    def set_gemm_sm_limit(self, n: int):
        self._rc(self.lib.stcat_set_gemm_sm_limit(int(n)), "set_gemm_sm_limit")

    def set_sm_cap(self, n: int):
        """SMs the persistent kernels size their grids for (0 = all); see include/stcat_b200.h stcat_set_sm_cap."""
        self._rc(self.lib.stcat_set_sm_cap(int(n)), "set_sm_cap")

    def attn_counts(self):
        """launches of the attention entry points per kernel family (stcat_debug_attn_counts)"""
        buf = (ctypes.c_longlong * 5)()
        self._rc(self.lib.stcat_debug_attn_counts(buf), "debug_attn_counts")
        return dict(zip(("sq", "tc", "mma", "small", "simt"), (int(x) for x in buf)))

    @staticmethod
    def _mat(t: torch.Tensor, name: str):
        if not t.is_cuda:
            raise StcatError(f"{name}: tensor is on {t.device}; the STCAT hot path has no CPU fallback")
        if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
            raise StcatError(f"{name}: expected a 2-D row-major view, got shape {tuple(t.shape)} stride {t.stride()}")
        ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
        return t.data_ptr(), ld, _dt(t)

    @staticmethod
    def _flat(t: Optional[torch.Tensor], name: str, dtype=None):
        if t is None:
            return None
        if not t.is_cuda or not t.is_contiguous():
            raise StcatError(f"{name}: must be a contiguous CUDA tensor")
        if dtype is not None and t.dtype != dtype:
            raise StcatError(f"{name}: expected {dtype}, got {t.dtype}")
        return t.data_ptr()

    # -- linear --------------------------------------------------------
    def linear_fwd(self, x, w, bias, y, relu=False, accumulate=False):
        (xp, ldx, xd), (wp, ldw, wd), (yp, ldy, yd) = self._mat(x, "x"), self._mat(w, "w"), self._mat(y, "y")
        M, K = x.shape
        N = w.shape[0]
        assert w.shape[1] == K and tuple(y.shape) == (M, N)
        self._rc(self.lib.stcat_linear_fwd(xp, ldx, xd, wp, ldw, wd, self._flat(bias, "bias", torch.float32), yp, ldy,
                                           yd, M, N, K, int(relu), int(accumulate), self._stream()), "linear_fwd")
        self.launches += 1

    def linear_dropout_fwd(self, x, w, bias, y, relu, drop):
        """y = drop(act(x w^T + bias)), the mask drawn in the GEMM epilogue; ``drop`` = (p, seed, offset), y contiguous"""
        (xp, ldx, xd), (wp, ldw, wd), (yp, ldy, yd) = self._mat(x, "x"), self._mat(w, "w"), self._mat(y, "y")
        M, K = x.shape
        N = w.shape[0]
        assert w.shape[1] == K and tuple(y.shape) == (M, N)
        self._rc(self.lib.stcat_linear_dropout_fwd(xp, ldx, xd, wp, ldw, wd, self._flat(bias, "bias", torch.float32), yp, ldy, yd,
                                                   M, N, K, int(relu), float(drop[0]), int(drop[1]), int(drop[2]),
                                                   self._stream()), "linear_dropout_fwd")
        self.launches += 1

    def linear_bwd_data(self, dy, w, dx, accumulate=False, relu_y=None, dbias=None, alpha=1.0):
        """dx = dy . w (+ dx); with relu_y ([M, K], the forward activation of the layer below) the result is zeroed
        where relu_y <= 0, and with dbias ([K] fp32) the column sums of the stored dx are accumulated into it.
        ``alpha`` != 1 (with relu_y, no accumulate): dx is scaled by it -- the 1 / keep factor of a dropout behind the ReLU."""
        (gp, ldg, gd), (wp, ldw, wd), (xp, ldx, xd) = self._mat(dy, "dy"), self._mat(w, "w"), self._mat(dx, "dx")
        M, N = dy.shape
        K = w.shape[1]
        assert w.shape[0] == N and tuple(dx.shape) == (M, K)
        yp, ldy, yd = (None, 0, F32) if relu_y is None else self._mat(relu_y, "relu_y")
        assert relu_y is None or tuple(relu_y.shape) == (M, K)
        if alpha != 1.0:
            assert relu_y is not None and not accumulate
            self._rc(self.lib.stcat_linear_bwd_data_scaled(gp, ldg, gd, wp, ldw, wd, xp, ldx, xd, yp, ldy, yd,
                                                           self._flat(dbias, "dbias", torch.float32), M, N, K, float(alpha),
                                                           self._stream()), "linear_bwd_data_scaled")
        else:
            self._rc(self.lib.stcat_linear_bwd_data(gp, ldg, gd, wp, ldw, wd, xp, ldx, xd, yp, ldy, yd,
                                                    self._flat(dbias, "dbias", torch.float32), M, N, K, int(accumulate),
                                                    self._stream()), "linear_bwd_data")
        self.launches += 1

    def linear_bwd_weight(self, dy, x, dw, db, accumulate=False):
        (gp, ldg, gd), (xp, ldx, xd), (wp, ldw, wd) = self._mat(dy, "dy"), self._mat(x, "x"), self._mat(dw, "dw")
        M, N = dy.shape
        K = x.shape[1]
        assert x.shape[0] == M and tuple(dw.shape) == (N, K) and wd == F32
        self._rc(self.lib.stcat_linear_bwd_weight(gp, ldg, gd, xp, ldx, xd, wp, ldw, self._flat(db, "db", torch.float32),
                                                  M, N, K, int(accumulate), self._stream()), "linear_bwd_weight")
        self.launches += 1 + (db is not None)

    def linear_group(self, kind: int, jobs):
        """One launch for up to 12 independent multi-term Linear GEMMs (include/stcat_b200.h, stcat_linear_group).
        kind 0 fwd / 1 bwd_data / 2 bwd_weight; each job is a dict with ``terms`` = [(a, b, bias-or-None), ...]
        (meaning per kind as in the header), ``out``, and optional ``relu``, ``accumulate``, ``dbias``."""
        assert 1 <= len(jobs) <= MAX_GROUP_JOBS
        arr = (_Job * len(jobs))()
        in_dt = None
        for j, job in enumerate(jobs):
            J = arr[j]
            terms = job["terms"]
            assert 1 <= len(terms) <= 3
            op, ldo, od = self._mat(job["out"], "out")
            rows, cols = job["out"].shape
            for t, (a, b, bias) in enumerate(terms):
                (ap, lda, ad), (bp, ldb, bd) = self._mat(a, "a"), self._mat(b, "b")
                assert ad == bd and (in_dt is None or in_dt == ad), "operands of a grouped launch share one dtype"
                in_dt = ad
                if kind == 0:
                    k = a.shape[1]
                    assert a.shape[0] == rows and tuple(b.shape) == (cols, k)
                elif kind == 1:
                    k = a.shape[1]
                    assert a.shape[0] == rows and tuple(b.shape) == (k, cols)
                else:
                    k = a.shape[0]
                    assert tuple(a.shape) == (k, rows) and tuple(b.shape) == (k, cols)
                T = J.term[t]
                T.a, T.lda, T.b, T.ldb, T.k = ap, lda, bp, ldb, k
                T.bias = self._flat(bias, "bias", torch.float32) if bias is not None else None
            J.nterms, J.rows, J.cols = len(terms), rows, cols
            J.relu, J.accumulate, J.out_dtype = int(job.get("relu", False)), int(job.get("accumulate", False)), od
            J.out, J.ldo = op, ldo
            db = job.get("dbias")
            J.dbias = self._flat(db, "dbias", torch.float32) if db is not None else None
        self._rc(self.lib.stcat_linear_group(kind, in_dt, arr, len(jobs), self._stream()), "linear_group")
        self.launches += 1 + (kind == 2 and any(j.get("dbias") is not None for j in jobs))

    # -- layernorm -----------------------------------------------------
    def layernorm_fwd(self, x, res, gamma, beta, y, y_bf16, mean, rstd, eps=1e-5, drop=None):
        """``drop`` = (p, seed, offset): train-mode dropout on x folded into the kernel (stcat_layernorm_dropout_fwd)"""
        rows, d = x.shape
        f = torch.float32
        args = (self._flat(x, "x", f), self._flat(res, "res", f), self._flat(gamma, "gamma", f), self._flat(beta, "beta", f),
                self._flat(y, "y", f), self._flat(y_bf16, "y_bf16", torch.bfloat16), self._flat(mean, "mean", f),
                self._flat(rstd, "rstd", f), rows, d, float(eps))
        if drop is not None and drop[0] > 0:
            self._rc(self.lib.stcat_layernorm_dropout_fwd(*args, float(drop[0]), int(drop[1]), int(drop[2]), self._stream()),
                     "layernorm_dropout_fwd")
        else:
            self._rc(self.lib.stcat_layernorm_fwd(*args, self._stream()), "layernorm_fwd")
        self.launches += 1

    def layernorm_bwd(self, dy, x, res, gamma, mean, rstd, dz, dgamma, dbeta, dz_bf16=None, dbias=None, drop=None):
        rows, d = x.shape
        f = torch.float32
        args = (self._flat(dy, "dy", f), self._flat(x, "x", f), self._flat(res, "res", f), self._flat(gamma, "gamma", f),
                self._flat(mean, "mean", f), self._flat(rstd, "rstd", f), self._flat(dz, "dz", f),
                self._flat(dz_bf16, "dz_bf16", torch.bfloat16), self._flat(dgamma, "dgamma", f), self._flat(dbeta, "dbeta", f),
                self._flat(dbias, "dbias", f), rows, d)
        if drop is not None and drop[0] > 0:
            self._rc(self.lib.stcat_layernorm_dropout_bwd(*args, float(drop[0]), int(drop[1]), int(drop[2]), self._stream()),
                     "layernorm_dropout_bwd")
        else:
            self._rc(self.lib.stcat_layernorm_bwd(*args, self._stream()), "layernorm_bwd")
        self.launches += 1

    # -- attention -----------------------------------------------------
    def dropout(self, x, out, p, seed, offset):
        """out = dropout mask(seed, offset) applied to x (same call on the gradient = backward); in place allowed"""
        assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype and x.shape == out.shape
        self._rc(self.lib.stcat_dropout(self._flat(x, "x"), self._flat(out, "out"), _dt(x), x.numel(), float(p), int(seed), int(offset),
                                        self._stream()), "dropout")
        self.launches += 1

    def dropout_bits(self, bits, cols, p, seed, offset):
        """bits [rows, wpr] int32: the keep mask of a [rows, cols] dropout site, one bit per element (stcat_dropout_bits)"""
        assert bits.dtype == torch.int32 and bits.is_contiguous() and bits.shape[1] * 32 >= cols
        self._rc(self.lib.stcat_dropout_bits(bits.data_ptr(), bits.shape[0], int(cols), bits.shape[1], float(p), int(seed), int(offset),
                                             self._stream()), "dropout_bits")
        self.launches += 1

    def _bits(self, bits, rows):
        if bits is None:
            return None, 0
        assert bits.is_cuda and bits.dtype == torch.int32 and bits.is_contiguous() and bits.shape[0] == rows
        return bits.data_ptr(), bits.shape[1]

    def attention_fwd(self, q1, q2, k1, k2, v, o, key_mask, lse, p_avg, B, H, Lq, Lk, scale, drop=None, bits=None):
        """q*: [B*Lq, H*32] views, k*/v: [B*Lk, H*32] views, o: [B*Lq, H*32]; q2/k2 share q1/k1's leading dim.
        ``bits``: optional precomputed keep bits of the dropout site (dropout_bits with rows = B*H*Lq, cols = Lk)."""
        (qp, ldq, qd) = self._mat(q1, "q1")
        (kp, ldk, _), (vp, ldv, _), (op, ldo, _) = self._mat(k1, "k1"), self._mat(v, "v"), self._mat(o, "o")
        q2p = k2p = None
        if q2 is not None:
            q2p, ldq2, _ = self._mat(q2, "q2")
            k2p, ldk2, _ = self._mat(k2, "k2")
            assert ldq2 == ldq and ldk2 == ldk
        assert q1.shape == (B * Lq, H * 32) and k1.shape == (B * Lk, H * 32) and v.shape == (B * Lk, H * 32)
        if drop is not None and drop[0] > 0:  # (p, seed, offset): dropout on the probabilities
            self._rc(self.lib.stcat_attention_dropout_fwd(
                qp, q2p, ldq, kp, k2p, ldk, vp, ldv, op, ldo, qd, self._flat(key_mask, "key_mask", torch.uint8),
                self._flat(lse, "lse", torch.float32), self._flat(p_avg, "p_avg", torch.float32), B, H, Lq, Lk, 32, float(scale),
                float(drop[0]), int(drop[1]), int(drop[2]), *self._bits(bits, B * H * Lq), self._stream()), "attention_dropout_fwd")
            self.launches += 1
            return
        self._rc(self.lib.stcat_attention_fwd(qp, q2p, ldq, kp, k2p, ldk, vp, ldv, op, ldo, qd,
                                              self._flat(key_mask, "key_mask", torch.uint8), self._flat(lse, "lse", torch.float32),
                                              self._flat(p_avg, "p_avg", torch.float32), B, H, Lq, Lk, 32, float(scale),
                                              self._stream()), "attention_fwd")
        self.launches += 1

    def attention_bwd(self, q1, q2, k1, k2, v, d_o, key_mask, lse, dp_avg, delta, dq1, dq2, dk1, dk2, dv, B, H, Lq, Lk,
                      scale, o=None, drop=None, bits=None):
        (qp, ldq, qd) = self._mat(q1, "q1")
        (kp, ldk, _), (vp, ldv, _), (gp, ldg, _) = self._mat(k1, "k1"), self._mat(v, "v"), self._mat(d_o, "d_o")
        (dqp, lddq, _), (dkp, lddk, _), (dvp, lddv, _) = self._mat(dq1, "dq1"), self._mat(dk1, "dk1"), self._mat(dv, "dv")
        q2p = k2p = dq2p = dk2p = None
        if q2 is not None:
            q2p, ldq2, _ = self._mat(q2, "q2")
            k2p, ldk2, _ = self._mat(k2, "k2")
            dq2p, lddq2, _ = self._mat(dq2, "dq2")
            dk2p, lddk2, _ = self._mat(dk2, "dk2")
            assert ldq2 == ldq and ldk2 == ldk and lddq2 == lddq and lddk2 == lddk
        op, ldo = (None, 0) if o is None else self._mat(o, "o")[:2]
        if drop is not None and drop[0] > 0:
            self._rc(self.lib.stcat_attention_dropout_bwd(
                qp, q2p, ldq, kp, k2p, ldk, vp, ldv, op, ldo, gp, ldg, qd, self._flat(key_mask, "key_mask", torch.uint8),
                self._flat(lse, "lse", torch.float32), self._flat(dp_avg, "dp_avg", torch.float32),
                self._flat(delta, "delta", torch.float32), dqp, dq2p, lddq, dkp, dk2p, lddk, dvp, lddv, B, H, Lq, Lk, 32,
                float(scale), float(drop[0]), int(drop[1]), int(drop[2]), *self._bits(bits, B * H * Lq), self._stream()),
                "attention_dropout_bwd")
            self.launches += 2
            return
        self._rc(self.lib.stcat_attention_bwd(qp, q2p, ldq, kp, k2p, ldk, vp, ldv, op, ldo, gp, ldg, qd,
                                              self._flat(key_mask, "key_mask", torch.uint8), self._flat(lse, "lse", torch.float32),
                                              self._flat(dp_avg, "dp_avg", torch.float32), self._flat(delta, "delta", torch.float32),
                                              dqp, dq2p, lddq, dkp, dk2p, lddk, dvp, lddv, B, H, Lq, Lk, 32, float(scale),
                                              self._stream()), "attention_bwd")
        self.launches += 2

    # -- element-wise --------------------------------------------------
    def add(self, a, b, out, out_bf16=None):
        f = torch.float32
        self._rc(self.lib.stcat_add(self._flat(a, "a", f), self._flat(b, "b", f), self._flat(out, "out", f),
                                    self._flat(out_bf16, "out_bf16", torch.bfloat16), a.numel(), self._stream()), "add")
        self.launches += 1

    def relu_bwd(self, y, dy):
        self._rc(self.lib.stcat_relu_bwd(self._flat(y, "y"), _dt(y), self._flat(dy, "dy"), _dt(dy), y.numel(),
                                         self._stream()), "relu_bwd")
        self.launches += 1

    def cast_bf16(self, x, out, transpose=False):
        rows, cols = x.shape
        self._rc(self.lib.stcat_cast_bf16(self._flat(x, "x", torch.float32), self._flat(out, "out", torch.bfloat16), rows,
                                          cols, int(transpose), self._stream()), "cast_bf16")
        self.launches += 1

    # -- anchor glue ---------------------------------------------------
    def pos_sine(self, mask, out, num_pos_feats: int, temperature: float, scale: float):
        """mask [n, H, W] uint8 -> out [n, H, W, 2F] fp32 (channels-last)"""
        n, H, W = mask.shape
        assert out.shape == (n, H, W, 2 * num_pos_feats)
        self._rc(self.lib.stcat_pos_sine(self._flat(mask, "mask", torch.uint8), self._flat(out, "out", torch.float32), n, H, W,
                                         int(num_pos_feats), float(temperature), float(scale), self._stream()), "pos_sine")
        self.launches += 1

    def box_interp(self, frame_ids, boxes, out, first: int):
        """frame_ids [m] int64 ascending, boxes [m, 4] fp32 -> out [n_frames, 4] fp32"""
        self._rc(self.lib.stcat_box_interp(self._flat(frame_ids, "frame_ids", torch.int64), self._flat(boxes, "boxes", torch.float32),
                                           int(frame_ids.shape[0]), self._flat(out, "out", torch.float32), int(first), int(out.shape[0]),
                                           self._stream()), "box_interp")
        self.launches += 1

    def anchor_sine_fwd(self, anchor, out, out_bf16=None):
        f = torch.float32
        self._rc(self.lib.stcat_anchor_sine_fwd(self._flat(anchor, "anchor", f), self._flat(out, "out", f),
                                                self._flat(out_bf16, "out_bf16", torch.bfloat16), anchor.shape[0], self._stream()),
                 "anchor_sine_fwd")
        self.launches += 1

    def anchor_sine_bwd(self, anchor, dy, danchor):
        f = torch.float32
        self._rc(self.lib.stcat_anchor_sine_bwd(self._flat(anchor, "anchor", f), self._flat(dy, "dy", f), self._flat(danchor, "danchor", f),
                                                anchor.shape[0], self._stream()), "anchor_sine_bwd")
        self.launches += 1

    def box_refine_fwd(self, delta, anchor, out, eps=1e-3):
        f = torch.float32
        self._rc(self.lib.stcat_box_refine_fwd(self._flat(delta, "delta", f), self._flat(anchor, "anchor", f), self._flat(out, "out", f),
                                               delta.numel(), float(eps), self._stream()), "box_refine_fwd")
        self.launches += 1

    def box_refine_bwd(self, out, anchor, g, ddelta, danchor, eps=1e-3):
        f = torch.float32
        self._rc(self.lib.stcat_box_refine_bwd(self._flat(out, "out", f), self._flat(anchor, "anchor", f), self._flat(g, "g", f),
                                               self._flat(ddelta, "ddelta", f), self._flat(danchor, "danchor", f), out.numel(),
                                               float(eps), self._stream()), "box_refine_bwd")
        self.launches += 1

    # -- loss ----------------------------------------------------------
    def stg_loss(self, coord, sted, act, attn, plan, coef5, losses, d_coord, d_sted, d_act, d_attn):
        """VideoSTGLoss values + gradients for all layers in one launch (include/stcat_b200.h, stcat_stg_loss).  ``plan`` carries
        the precomputed target tensors (loss.STGLossPlan)."""
        f = torch.float32
        nl, n, _ = coord.shape
        b, t = plan.b, plan.t
        carr = (c_float * 5)(*[float(c) for c in coef5])
        self._rc(self.lib.stcat_stg_loss(
            self._flat(coord, "coord", f), self._flat(sted, "sted", f), self._flat(act, "act", f), self._flat(attn, "attn", f),
            self._flat(plan.slice, "slice", torch.int64), self._flat(plan.target_boxes, "target_boxes", f),
            self._flat(plan.time_mask_u8, "time_mask", torch.uint8), self._flat(plan.distrib, "distrib", f),
            self._flat(plan.neg_f, "neg_f", f), self._flat(plan.nb_neg, "nb_neg", f), self._flat(plan.bce_weight, "bce_weight", f),
            self._flat(plan.actioness, "actioness", f), carr, float(plan.num_boxes), nl, n, b, t, int(plan.slice.numel()),
            self._flat(losses, "losses", f), self._flat(d_coord, "d_coord", f), self._flat(d_sted, "d_sted", f),
            self._flat(d_act, "d_act", f), self._flat(d_attn, "d_attn", f), self._stream()), "stg_loss")
        self.launches += 1

    # -- optimizer-side step ---------------------------------------------
    def sumsq(self, x, accum):
        """accum (1-element fp32 device tensor) += sum x^2"""
        f = torch.float32
        self._rc(self.lib.stcat_sumsq(self._flat(x, "x", f), x.numel(), self._flat(accum, "accum", f), self._stream()), "sumsq")
        self.launches += 1

    def adamw_step(self, p, g, m, v, ema, shadow, lr, beta1, beta2, eps, weight_decay, step, total_sumsq, max_norm, ema_decay):
        """AdamW (+ global-norm clipping, EMA, bf16 shadow refresh) on one contiguous fp32 range; include/stcat_b200.h"""
        f = torch.float32
        n = p.numel()
        assert g.numel() == n and m.numel() == n and v.numel() == n and (ema is None or ema.numel() == n)
        assert shadow is None or shadow.numel() == n
        self._rc(self.lib.stcat_adamw_step(
            self._flat(p, "p", f), self._flat(g, "g", f), self._flat(m, "m", f), self._flat(v, "v", f), self._flat(ema, "ema", f),
            self._flat(shadow, "shadow", torch.bfloat16), n, float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
            int(step), self._flat(total_sumsq, "total_sumsq", f), float(max_norm), float(ema_decay), self._stream()), "adamw_step")
        self.launches += 1

    # -- post-process / map2d -------------------------------------------
    def sted_score(self, sted, durations, score, best):
        b, t, _ = sted.shape
        self._rc(self.lib.stcat_sted_score(self._flat(sted, "sted", torch.float32), self._flat(durations, "durations", torch.int32),
                                           self._flat(score, "score", torch.float32), self._flat(best, "best", torch.int32), b, t,
                                           self._stream()), "sted_score")
        self.launches += 1

    def map2d_pool(self, x, valid, out):
        B, N, d = x.shape
        self._rc(self.lib.stcat_map2d_pool(self._flat(x, "x", torch.float32), self._flat(valid, "valid", torch.uint8),
                                           self._flat(out, "map", torch.float32), B, N, d, self._stream()), "map2d_pool")
        self.launches += 1

    # -- layout glue (csrc/assembly.cu) ----------------------------------
    def token_assembly(self, vis, vpos, text, f2v, frame_cls, local_pos, X, POS, qk_op, x_op):
        """vis / vpos [n, d, H, W] fp32, text [L, b, d] fp32 -> X / POS [n, S, d] fp32 (+ bf16 operand copies or None)"""
        f, bf = torch.float32, torch.bfloat16
        n, d = vis.shape[0], vis.shape[1]
        HW = vis[0, 0].numel()
        L, b = text.shape[0], text.shape[1]
        assert tuple(X.shape) == (n, 1 + HW + L, d) and X.shape == POS.shape and vpos.shape == vis.shape
        self._rc(self.lib.stcat_token_assembly(
            self._flat(vis, "vis", f), self._flat(vpos, "vpos", f), self._flat(text, "text", f), self._flat(f2v, "f2v", torch.int64),
            self._flat(frame_cls, "frame_cls", f), self._flat(local_pos, "local_pos", f), self._flat(X, "X", f), self._flat(POS, "POS", f),
            self._flat(qk_op, "qk_op", bf), self._flat(x_op, "x_op", bf), n, d, HW, L, b, self._stream()), "token_assembly")
        self.launches += 1

    def token_assembly_bwd(self, dX, dvis, dtext, dcls, vid_start, HW, L, b):
        f = torch.float32
        n, S, d = dX.shape
        assert S == 1 + HW + L
        self._rc(self.lib.stcat_token_assembly_bwd(
            self._flat(dX, "dX", f), self._flat(dvis, "dvis", f), self._flat(dtext, "dtext", f), self._flat(dcls, "dcls", f),
            self._flat(vid_start, "vid_start", torch.int64), n, d, HW, L, b, self._stream()), "token_assembly_bwd")
        self.launches += 1

    def mem_operands(self, X, POS, mem_op, pos_op, mempos_op, cls):
        """X / POS [n, S, d] fp32 -> bf16 [n (S-1), d] operands of rows 1.. and the fp32 CLS rows [n, d]"""
        f, bf = torch.float32, torch.bfloat16
        n, S, d = X.shape
        self._rc(self.lib.stcat_mem_operands(
            self._flat(X, "X", f), self._flat(POS, "POS", f), self._flat(mem_op, "mem_op", bf), self._flat(pos_op, "pos_op", bf),
            self._flat(mempos_op, "mempos_op", bf), self._flat(cls, "cls", f), n, S, d, self._stream()), "mem_operands")
        self.launches += 1

    def mem_operands_bwd(self, g_mem, g_mempos, g_cls, dX):
        n, S, d = dX.shape
        self._rc(self.lib.stcat_mem_operands_bwd(
            self._flat(g_mem, "g_mem"), F32 if g_mem is None else _dt(g_mem), self._flat(g_mempos, "g_mempos"),
            F32 if g_mempos is None else _dt(g_mempos), self._flat(g_cls, "g_cls", torch.float32), self._flat(dX, "dX", torch.float32),
            n, S, d, self._stream()), "mem_operands_bwd")
        self.launches += 1

    def template_fwd(self, videos_cls, frames_cls, f2v, Wc, bc, Wg, bg, Wb, bb, Wa, ba, content, gamma, beta, mod_op, anchor, temp_query):
        f, bf = torch.float32, torch.bfloat16
        n, d = frames_cls.shape
        b, q = videos_cls.shape[0], Wa.shape[0]
        self._rc(self.lib.stcat_template_fwd(
            self._flat(videos_cls, "videos_cls", f), self._flat(frames_cls, "frames_cls", f), self._flat(f2v, "f2v", torch.int64),
            self._flat(Wc, "Wc", bf), self._flat(bc, "bc", f), self._flat(Wg, "Wg", bf), self._flat(bg, "bg", f),
            self._flat(Wb, "Wb", bf), self._flat(bb, "bb", f), self._flat(Wa, "Wa", bf), self._flat(ba, "ba", f),
            self._flat(content, "content", f), self._flat(gamma, "gamma", f), self._flat(beta, "beta", f),
            self._flat(mod_op, "mod_op", bf), self._flat(anchor, "anchor", f), self._flat(temp_query, "temp_query", f), n, b, d, q,
            self._stream()), "template_fwd")
        self.launches += 2

    def template_bwd(self, g_anchor, g_temp, anchor, videos_cls, frames_cls, f2v, vid_start, gamma, beta, mod_op, Wc, Wg, Wb, Wa,
                     dpq_op, dmod, dpre, d_frames_cls, d_videos_cls, dWc, dbc, dWg, dbg, dWb, dbb, dWa, dba):
        f, bf, i64 = torch.float32, torch.bfloat16, torch.int64
        n, d = frames_cls.shape
        b, q = videos_cls.shape[0], Wa.shape[0]
        F_ = lambda t, nm: self._flat(t, nm, f)
        self._rc(self.lib.stcat_template_bwd(
            F_(g_anchor, "g_anchor"), F_(g_temp, "g_temp"), F_(anchor, "anchor"), F_(videos_cls, "videos_cls"),
            F_(frames_cls, "frames_cls"), self._flat(f2v, "f2v", i64), self._flat(vid_start, "vid_start", i64), F_(gamma, "gamma"),
            F_(beta, "beta"), self._flat(mod_op, "mod_op", bf), self._flat(Wc, "Wc", bf), self._flat(Wg, "Wg", bf),
            self._flat(Wb, "Wb", bf), self._flat(Wa, "Wa", bf), self._flat(dpq_op, "dpq_op", bf), F_(dmod, "dmod"), F_(dpre, "dpre"),
            F_(d_frames_cls, "d_frames_cls"), F_(d_videos_cls, "d_videos_cls"), F_(dWc, "dWc"), F_(dbc, "dbc"), F_(dWg, "dWg"),
            F_(dbg, "dbg"), F_(dWb, "dWb"), F_(dbb, "dbb"), F_(dWa, "dWa"), F_(dba, "dba"), n, b, d, q, self._stream()), "template_bwd")
        self.launches += 3

    def box_head_fwd(self, h, W, bias, anchor, out, sine, sine_op, eps=1e-3):
        """h bf16 [R, K] (row-major view), W bf16 [4, K], anchor / out fp32 [R, 4]; sine fp32 [R, 512] + bf16 copy, or None"""
        hp, ldh, hdt = self._mat(h, "h")
        assert hdt == BF16 and W.shape[0] == 4 and W.shape[1] == h.shape[1]
        R, K = h.shape
        self._rc(self.lib.stcat_box_head_fwd(hp, ldh, self._flat(W, "W", torch.bfloat16), self._flat(bias, "bias", torch.float32),
                                             self._flat(anchor, "anchor", torch.float32), self._flat(out, "out", torch.float32),
                                             self._flat(sine, "sine", torch.float32), self._flat(sine_op, "sine_op", torch.bfloat16),
                                             R, K, float(eps), self._stream()), "box_head_fwd")
        self.launches += 1

    def mul_cast(self, a, b, out_f32, out_bf16, c_in=None, c_out=None):
        """out_f32 (or None) = a[:, :c] * b and out_bf16 [R, c] its bf16 copy; a fp32 [R, >= c] row-major view, b fp32 [R, c]"""
        ap, lda, adt = self._mat(a, "a")
        assert adt == F32
        R, c = b.shape
        self._rc(self.lib.stcat_mul_cast(ap, lda, self._flat(b, "b", torch.float32), self._flat(out_f32, "out_f32", torch.float32),
                                         self._flat(out_bf16, "out", torch.bfloat16), self._flat(c_in, "c_in", torch.float32),
                                         self._flat(c_out, "c_out", torch.bfloat16), R, c,
                                         self._stream()), "mul_cast")
        self.launches += 1

    def mul_cast_bwd(self, g, a, db):
        ap, lda, adt = self._mat(a, "a")
        assert adt == F32
        R, c = db.shape
        self._rc(self.lib.stcat_mul_cast_bwd(self._flat(g, "g"), _dt(g), ap, lda, self._flat(db, "db", torch.float32), R, c,
                                             self._stream()), "mul_cast_bwd")
        self.launches += 1

    def box_head_bwd(self, g, out, anchor, W, h, dd_op, dh, danchor, eps=1e-3):
        f, bf = torch.float32, torch.bfloat16
        hp, ldh, hdt = self._mat(h, "h")
        assert hdt == BF16
        R, K = h.shape
        self._rc(self.lib.stcat_box_head_bwd(self._flat(g, "g", f), self._flat(out, "out", f), self._flat(anchor, "anchor", f),
                                             self._flat(W, "W", bf), hp, ldh, self._flat(dd_op, "dd_op", bf), self._flat(dh, "dh", bf),
                                             self._flat(danchor, "danchor", f), R, K, float(eps), self._stream()), "box_head_bwd")
        self.launches += 1

    def cls_gather(self, X, video, pos, Y, qk_op, y_op, r):
        """X fp32 [n, S, d], video fp32 [1, d], pos fp32 [1 + n, d] or None -> Y fp32 [1 + n, d] (+ bf16 operands or None)"""
        f, bf = torch.float32, torch.bfloat16
        n, S, d = X.shape
        self._rc(self.lib.stcat_cls_gather(self._flat(X, "X", f), self._flat(video, "video", f), self._flat(pos, "pos", f),
                                           self._flat(Y, "Y", f), self._flat(qk_op, "qk_op", bf), self._flat(y_op, "y_op", bf), n, S,
                                           int(r), d, self._stream()), "cls_gather")
        self.launches += 1

    def cls_scatter(self, Y, X, X_op, r, qk_next=None, pos=None):
        f, bf = torch.float32, torch.bfloat16
        n, S, d = X.shape
        self._rc(self.lib.stcat_cls_scatter(self._flat(Y, "Y", f), self._flat(X, "X", f), self._flat(X_op, "X_op", bf),
                                            self._flat(qk_next, "qk_next", bf), self._flat(pos, "pos", f), n, S, int(r), d,
                                            self._stream()), "cls_scatter")
        self.launches += 1

"""autograd operators of the hot path: each one is a forward/backward pair of C-ABI kernels.

These replace the torch.nn calls the reference makes (nn.Linear, nn.LayerNorm, the bmm/softmax/bmm
core of nn.MultiheadAttention and of the reference's own attention.py).  They are composed by
``encoder.py`` / ``decoder.py`` into modules with the reference's interface.

Precision:
  * ``"fp32"``: exact-fp32 SIMT kernels; gated <= 1e-3 (measured ~1e-5) against the fp32 reference.
  * ``"bf16"``: GEMM / attention operands rounded to bf16 (tcgen05 tensor cores, fp32 accumulation),
    fp32 residual stream, LayerNorm and softmax statistics.
"""
from __future__ import annotations

import os
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

_backend = None
_precision = "fp32"
_fuse_grads = False


def set_grad_fusion(on: bool):
    """When on, weight / bias / LayerNorm-affine gradients are accumulated by the wgrad kernels straight into
    ``param.grad`` (which must already exist, e.g. as views of one flat buffer) and autograd is handed ``None``:
    no per-parameter temporaries, fills or ``grad += dw`` kernels (about 700 tiny launches per step).  Off by
    default because it bypasses AccumulateGrad hooks (torch DDP relies on them); ``bench.py`` turns it on and
    all-reduces the flat buffer itself."""
    global _fuse_grads
    _fuse_grads = bool(on)


# ---- leaf streams --------------------------------------------------------------------------------------------------------
# Weight / bias gradients are leaves of the backward pass: nothing downstream waits for them until the gradient exchange /
# the optimizer.  With ``set_leaf_streams(True)`` (and in-place accumulation into .grad, ``set_grad_fusion``) the wgrad GEMMs
# and bias column sums of an autograd node are issued on a side stream forked from the node's stream, so the dependent chain
# of a decoder / temporal layer's backward (data gradients, attention, LayerNorm) no longer queues behind ~270 small leaf
# kernels per step.  ``join_leaf_streams()`` makes the current stream wait for all of them: ``dp.GradSync`` (before a range is
# reduced), ``optim.FusedAdamW.step`` and ``dp.FlatGrads.zero`` call it; call it yourself before reading ``.grad`` otherwise.
_leaf = {"on": False, "streams": {}, "pending": [], "max_rows": 1 << 30}
_LEAF_SMS = int(os.environ.get("STCAT_LEAF_SMS", "0"))  # experiment: SM cap of the leaf-stream GEMMs (0 = none)


def set_leaf_streams(flag: bool, max_rows: int = 1 << 30):
    """``max_rows``: only nodes whose operand has at most this many rows are off-loaded."""
    _leaf["on"] = bool(flag)
    _leaf["max_rows"] = int(max_rows)


def leaf_streams():
    """leaf streams with work issued since the last join (only these may be waited for: under CUDA-graph capture a stream that
    was not forked from the capturing stream must not be touched)"""
    return list(_leaf["pending"])


def join_leaf_streams(stream=None):
    if not _leaf["pending"]:
        return
    cur = stream if stream is not None else torch.cuda.current_stream()
    for st in _leaf["pending"]:
        cur.wait_stream(st)
    _leaf["pending"] = []


class _on_leaf:
    """``with _on_leaf(t0, t1, ...):`` -- the launches inside read the given tensors and accumulate into .grad buffers; they
    run on the leaf stream of the current stream (no-op when leaf streams are off or on the CPU)."""

    def __init__(self, *tensors):
        self.tensors = [t for t in tensors if t is not None]
        self.ctx = None

    def __enter__(self):
        if not _leaf["on"] or not self.tensors or not self.tensors[0].is_cuda or self.tensors[0].shape[0] > _leaf["max_rows"]:
            return self
        cur = torch.cuda.current_stream()
        side = _leaf["streams"].get(cur.cuda_stream)
        if side is None:
            side = _leaf["streams"][cur.cuda_stream] = torch.cuda.Stream(self.tensors[0].device)
        side.wait_stream(cur)
        if side not in _leaf["pending"]:
            _leaf["pending"].append(side)
        for t in self.tensors:
            t.record_stream(side)
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        if _LEAF_SMS:
            get_backend().set_gemm_sm_limit(_LEAF_SMS)
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            if _LEAF_SMS:
                get_backend().set_gemm_sm_limit(0)
            self.ctx.__exit__(*exc)
            self.ctx = None


def _wgrad(be, dyo, xo, W, b, need_w, need_b, r0=None, r1=None):
    """dW (+)= dy^T x, db (+)= colsum(dy) for rows [r0, r1) of W / b.  Returns (dw, db) for autograd, or Nones when
    the result was accumulated in place into W.grad / b.grad."""
    if not need_w:
        return None, None
    want_b = b is not None and need_b
    if _fuse_grads and W.grad is not None and (not want_b or b.grad is not None):
        gw = W.grad if r0 is None else W.grad[r0:r1]
        gb = None if not want_b else (b.grad if r0 is None else b.grad[r0:r1])
        with _on_leaf(dyo, xo):
            be.linear_bwd_weight(dyo, xo, gw, gb, accumulate=True)
        return None, None
    if r0 is None:
        dw = torch.empty(W.shape, dtype=torch.float32, device=dyo.device)
        db = torch.empty(b.shape, dtype=torch.float32, device=dyo.device) if want_b else None
        be.linear_bwd_weight(dyo, xo, dw, db)
        return dw, db
    dw = torch.zeros(W.shape, dtype=torch.float32, device=dyo.device)
    db = torch.zeros(b.shape, dtype=torch.float32, device=dyo.device) if want_b else None
    be.linear_bwd_weight(dyo, xo, dw[r0:r1], None if db is None else db[r0:r1])
    return dw, db


def _ln_bwd(be, dy, x, res, gamma, beta, mean, rstd, dz, dz_op=None, lin_bias=None, drop=None):
    """LayerNorm backward; returns (dgamma, dbeta, dbias) for autograd, each None when it was accumulated into
    .grad.  ``dz_op`` (bf16 [rows, d]) receives the GEMM-operand copy of dz; ``lin_bias`` is the bias parameter of
    the Linear whose output is ``x``: its gradient is colsum(dz), which the kernel sums anyway."""
    d = x.shape[1]
    fused = _fuse_grads and gamma.grad is not None and beta is not None and beta.grad is not None and \
        (lin_bias is None or lin_bias.grad is not None)
    if fused:
        be.layernorm_bwd(dy, x, res, gamma.detach(), mean, rstd, dz, gamma.grad, beta.grad, dz_bf16=dz_op,
                         dbias=None if lin_bias is None else lin_bias.grad, drop=drop)
        return None, None, None
    dg = torch.zeros(d, dtype=torch.float32, device=dy.device)
    dbt = torch.zeros(d, dtype=torch.float32, device=dy.device)
    dlb = None if lin_bias is None else torch.zeros(d, dtype=torch.float32, device=dy.device)
    be.layernorm_bwd(dy, x, res, gamma.detach(), mean, rstd, dz, dg, dbt, dz_bf16=dz_op, dbias=dlb, drop=drop)
    return dg, dbt, dlb


def get_backend():
    global _backend
    if _backend is None:
        from .cabi import CudaBackend

        _backend = CudaBackend()
    return _backend


def set_backend(b):
    """Install a backend object.  Product code never calls this; tests use it to drive the host-side
    composition with the torch-CPU emulation in tests/emu_backend.py."""
    global _backend
    _backend = b


def set_precision(p: str):
    global _precision
    if p not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {p!r}")
    _precision = p


def get_precision() -> str:
    return _precision


# bf16 shadows of the fp32 master weights, refreshed when the parameter changes (optimizer.step bumps
# ``_version``); keyed by address / shape so slices of a packed in_proj_weight get their own entry.  An address and a
# version do not identify a weight (a new Parameter can be allocated where a freed one lived): every entry also holds a
# weak reference to the storage object it was cast from and is dropped when that storage dies or is a different one.
# Writers that update the masters through raw pointers (optim.FusedAdamW without shadows) call clear_weight_cache().
_wcache = {}


def clear_weight_cache():
    _wcache.clear()


# optional: maps a weight (or a contiguous slice of one) to its bf16 shadow kept up to date by the fused optimizer step
# (optim.FusedAdamW writes the shadows in the same pass that updates the fp32 masters: no per-step cast kernels)
_shadow_provider = None


def set_shadow_provider(fn):
    global _shadow_provider
    _shadow_provider = fn
    _wcache.clear()


def _operand(t: torch.Tensor, is_weight: bool = False) -> torch.Tensor:
    """GEMM operand in the active precision (2-D, row-major)."""
    if _precision == "fp32" or t.dtype == torch.bfloat16:
        return t
    be = get_backend()
    if is_weight and _shadow_provider is not None:
        sh = _shadow_provider(t)
        if sh is not None:
            return sh
    if is_weight:
        key = (t.data_ptr(), tuple(t.shape), t.stride(0))
        storage = t.untyped_storage()
        hit = _wcache.get(key)
        if hit is not None and hit[0] == t._version and hit[2]() is storage:
            return hit[1]
    src = t if t.is_contiguous() else t.contiguous()
    out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    be.cast_bf16(src, out)
    if is_weight:
        if len(_wcache) > 4096:  # entries of dead storages (models that were freed) are not worth a sweep each call
            for k in [k for k, v in _wcache.items() if v[2]() is None]:
                del _wcache[k]
        _wcache[key] = (t._version, out, weakref.ref(storage))
    return out


def _rows(x: torch.Tensor, k: int) -> torch.Tensor:
    """[..., k] -> contiguous-row 2-D view [M, k] (copy only if rows are not unit-stride)."""
    x2 = x.reshape(-1, k)
    if x2.shape[0] > 0 and (x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < k)):
        x2 = x2.contiguous()
    return x2


# ---- train-mode dropout ---------------------------------------------------------------------------------------------
# Counter-based masks (include/stcat_b200.h, stcat_dropout): every dropout site of a forward pass reserves a range of the
# stream identified by (seed, offset); its backward regenerates the same mask from the saved pair.  The offset is host
# state, so a CUDA-graph replay would repeat the masks of the captured step: train with dropout eagerly (or re-capture).
_drop_state = {"seed": None, "offset": 0}


def set_dropout_seed(seed: int):
    """(Re)start the dropout stream (also resets the offset): same seed -> same masks, used by the tests."""
    _drop_state["seed"] = int(seed) & ((1 << 63) - 1)
    _drop_state["offset"] = 0


def _drop_scale(p: float) -> float:
    """1 / keep as the kernels compute it (csrc/common.cuh make_drop): 2^24 / (2^24 - floor(float32(p) * 2^24))"""
    t = min(int(float(torch.tensor(p, dtype=torch.float32)) * 16777216.0), 16777215)
    return float(torch.tensor(16777216.0 / (16777216.0 - t), dtype=torch.float32))


_bits_streams = {}


def _attn_keep_bits(be, B, H, L, drop, like):
    """Keep bits of an attention dropout site [B*H*L, L] for the tcgen05 kernels (include/stcat_b200.h stcat_dropout_bits): one
    word per 32 probabilities instead of a hash per probability inside the kernels' MUFU-bound loops.  Generated on a side
    stream forked here (the caller joins with the returned stream before the attention launch), so the generator runs under the
    projection GEMMs in front of the attention.  Returns (bits or None, side stream or None); None for shapes the tcgen05 kernels
    do not take (every other kernel hashes)."""
    if not (like.is_cuda and _opdtype() == torch.bfloat16 and 64 <= L <= 512 and (L > 128 or B * H >= 16)):
        return None, None
    bits = torch.empty(B * H * L, L // 32 + 2, dtype=torch.int32, device=like.device)
    cur = torch.cuda.current_stream()
    side = _bits_streams.get(cur.cuda_stream)
    if side is None:
        side = _bits_streams[cur.cuda_stream] = torch.cuda.Stream(like.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        bits.record_stream(side)
        be.dropout_bits(bits, L, *drop)
    return bits, side


def _drop_reserve(n: int):
    if _drop_state["seed"] is None:
        set_dropout_seed(torch.initial_seed())
    off = _drop_state["offset"]
    _drop_state["offset"] = off + int(n)
    return _drop_state["seed"], off


class DropoutFn(Function):
    """y = dropout(x, p) (train mode), mask regenerated in backward."""

    @staticmethod
    def forward(ctx, x, p: float):
        be = get_backend()
        xd = x.detach()
        xd = xd if xd.is_contiguous() else xd.contiguous()
        seed, off = _drop_reserve(xd.numel())
        out = torch.empty_like(xd)
        be.dropout(xd, out, p, seed, off)
        ctx.meta = (p, seed, off)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        p, seed, off = ctx.meta
        g = g if g.is_contiguous() else g.contiguous()
        dg = torch.empty_like(g)
        get_backend().dropout(g, dg, p, seed, off)
        return dg, None


def dropout(x, p: float):
    """Train-mode dropout site (callers pass p = 0 in eval mode)."""
    return x if not p else DropoutFn.apply(x, float(p))


# SM cap for the GEMMs of ops issued inside ``with sm_limit(n):`` -- and of their backward nodes, whenever autograd runs them
# (the cap is recorded on the node).  See include/stcat_b200.h stcat_set_gemm_sm_limit.
_sm_limit = 0


class sm_limit:
    def __init__(self, n: int):
        self.n = int(n)

    def __enter__(self):
        global _sm_limit
        self.prev, _sm_limit = _sm_limit, self.n

    def __exit__(self, *exc):
        global _sm_limit
        _sm_limit = self.prev


class _capped:
    """applies a node's SM cap to the backend for the launches inside"""

    def __init__(self, be, n):
        self.be, self.n = be, n

    def __enter__(self):
        if self.n:
            self.be.set_gemm_sm_limit(self.n)

    def __exit__(self, *exc):
        if self.n:
            self.be.set_gemm_sm_limit(0)


# Phase marks (measurement only, bench.py --phases): timing events recorded inside a captured step as EXTERNAL event nodes, so
# the phase boundaries of a graph replay can be read without a profiler attached.  None = off.
_phase_events = None


def enable_phase_marks(on: bool = True):
    global _phase_events
    _phase_events = {} if on else None


def phase_events():
    return _phase_events


def mark_phase(name: str):
    if _phase_events is None or not torch.cuda.is_available():
        return
    ev = _phase_events.get(name)
    if ev is None:
        ev = _phase_events[name] = torch.cuda.Event(enable_timing=True, external=True)
    ev.record()


_GRAD_SINK = os.environ.get("STCAT_GRAD_SINK", "1") != "0"


def set_grad_sink(on: bool):
    """Data gradients of the decoder's memory operands accumulate inside the GEMM launches (default on; see _GradSink)."""
    global _GRAD_SINK
    _GRAD_SINK = bool(on)


class _GradSink:
    """Accumulation target for the data gradient of a GEMM operand that MANY Linear nodes read (the encoder memory: 24
    memory-side projections of the two decoders).  autograd would sum their 24 [n M, d] gradients with 22 element-wise add
    kernels, issued on the consumer's stream -- measured as a ~75 us serial chain in front of the encoder's backward pass.
    A Linear node whose input carries a sink (attribute ``_stcat_sink``, set by ``mem_operands``) instead lets its
    data-gradient GEMM accumulate into the sink's buffer (``accumulate`` epilogue: one rounding per contribution instead of
    two) and returns None for that input; the producer of the operand (``MemOperandsFn.backward``), which the engine runs
    after every consumer taking part in this backward pass, collects the buffer.  No counting: a backward pass that reaches
    only some consumers just finds fewer contributions."""

    def __init__(self):
        self.buf = None
        self.event = None

    def accumulate(self, fn, shape, dtype, device):
        """fn(buf, accumulate: bool) launches the GEMM"""
        first = self.buf is None
        if first:
            self.buf = torch.empty(shape, dtype=dtype, device=device)
        fn(self.buf, not first)
        if self.buf.is_cuda:
            self.event = torch.cuda.current_stream().record_event()

    def take(self):
        """the accumulated gradient (or None), made safe to read on the current stream; the sink is empty afterwards"""
        buf, ev = self.buf, self.event
        self.buf = self.event = None
        if buf is not None and ev is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            buf.record_stream(cur)
        return buf


def _sink_of(x):
    return getattr(x, "_stcat_sink", None) if _GRAD_SINK else None


class LinearFn(Function):
    """y = act(x W[r0:r1]^T + b[r0:r1])  (r0/r1 = None: the whole parameter).  The row range lets the packed
    ``in_proj_weight`` of an MHA be used slice by slice without autograd slicing nodes."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool, out_bf16: bool, r0, r1, x_op=None):
        be = get_backend()
        wd = weight.detach()
        bd = None if bias is None else bias.detach()
        if r0 is not None:
            wd = wd[r0:r1]
            bd = None if bd is None else bd[r0:r1]
        N, K = wd.shape
        x2 = _rows(x.detach(), K)
        # x_op: the caller's GEMM-operand copy of x (made once for all consumers of x); gradients still flow to x
        xo = _rows(x_op.detach(), K) if (x_op is not None and _precision == "bf16") else _operand(x2)
        wo = _operand(wd, True)
        M = x2.shape[0]
        odt = torch.bfloat16 if (out_bf16 and _precision == "bf16") else torch.float32
        y = torch.empty(M, N, dtype=odt, device=x.device)
        ctx.sm_limit = _sm_limit
        with _capped(be, ctx.sm_limit):
            be.linear_fwd(xo, wo, bd, y, relu=relu)
        ctx.relu = relu
        ctx.x_shape = x.shape
        ctx.x_dtype = x.dtype
        ctx.rows = (r0, r1)
        ctx.sink = _sink_of(x) if x.dim() == 2 else None
        ctx.save_for_backward(xo, weight, bias, y if relu else None)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        be = get_backend()
        xo, weight, bias, y = ctx.saved_tensors
        r0, r1 = ctx.rows
        wd = weight.detach() if r0 is None else weight.detach()[r0:r1]
        N, K = wd.shape
        dy2 = _rows(dy, N)
        if ctx.relu:
            # A bf16 ReLU output is a hidden activation that feeds exactly one GEMM (callers pass out_bf16 only then): its gradient
            # was produced for this node alone by that GEMM's data-gradient launch, so the mask is applied in place.  Otherwise
            # (fp32 outputs may be user-visible) the incoming gradient is left untouched.
            if not (y.dtype == torch.bfloat16 and dy2.dtype == torch.bfloat16):
                dy2 = dy2.clone() if dy2.data_ptr() == dy.data_ptr() else dy2
            be.relu_bwd(y, dy2)
        dyo = _operand(dy2)
        M = dy2.shape[0]
        dx = None
        with _capped(be, ctx.sm_limit):
            if ctx.needs_input_grad[0] and ctx.sink is not None:
                ctx.sink.accumulate(lambda buf, acc: be.linear_bwd_data(dyo, _operand(wd, True), buf, accumulate=acc),
                                    (M, K), ctx.x_dtype, dy.device)
            elif ctx.needs_input_grad[0]:
                dx = torch.empty(M, K, dtype=ctx.x_dtype, device=dy.device)
                be.linear_bwd_data(dyo, _operand(wd, True), dx)
                dx = dx.view(ctx.x_shape)
            dw, db = _wgrad(be, dyo, xo, weight, bias, ctx.needs_input_grad[1], ctx.needs_input_grad[2], r0, r1)
        return dx, dw, db, None, None, None, None, None


def linear(x, weight, bias=None, relu: bool = False, out_bf16: bool = False, rows=None, x_op=None):
    r0, r1 = rows if rows is not None else (None, None)
    return LinearFn.apply(x, weight, bias, relu, out_bf16, r0, r1, x_op)


def operand_copy(x):
    """Non-differentiable GEMM-operand copy of ``x`` (bf16 in bf16 mode, None in fp32 mode), to be handed as
    ``x_op`` to every Linear that consumes ``x``: one cast kernel instead of one per consumer."""
    if _precision == "fp32":
        return None
    xd = x.detach()
    if xd.dtype == torch.bfloat16:
        return xd
    xd = xd if xd.is_contiguous() else xd.contiguous()
    out = torch.empty(xd.shape, dtype=torch.bfloat16, device=xd.device)
    get_backend().cast_bf16(xd.view(-1, xd.shape[-1]), out.view(-1, xd.shape[-1]))
    return out


class LinearSumFn(Function):
    """y = sum_i (x_i W_i^T + b_i): several Linear layers whose outputs the reference adds
    (query_decoder.py:329-339 q = q_content + q_time + q_pos; :355-366 k = k_content + k_pos), run as
    GEMMs accumulating into one output so the partial results never round-trip through HBM."""

    @staticmethod
    def forward(ctx, out_bf16, nterms, *args):
        be = get_backend()
        xs, ws, bs = args[:nterms], args[nterms:2 * nterms], args[2 * nterms:3 * nterms]
        x_ops = args[3 * nterms:4 * nterms]
        N, _ = ws[0].shape
        lead = xs[0].shape[:-1]
        odt = torch.bfloat16 if (out_bf16 and _precision == "bf16") else torch.float32
        y = None
        saved = []
        ctx.sm_limit = _sm_limit
        for i, (x, w, b) in enumerate(zip(xs, ws, bs)):
            K = w.shape[1]
            x2 = _rows(x.detach(), K)
            xo = _rows(x_ops[i].detach(), K) if (x_ops[i] is not None and _precision == "bf16") else _operand(x2)
            if y is None:
                y = torch.empty(x2.shape[0], N, dtype=odt, device=x.device)
            with _capped(be, ctx.sm_limit):
                be.linear_fwd(xo, _operand(w.detach(), True), None if b is None else b.detach(), y, accumulate=i > 0)
            saved += [xo, w]
        ctx.nterms = nterms
        ctx.biases = bs
        ctx.sinks = [(_sink_of(x) if x.dim() == 2 else None) for x in xs]
        ctx.meta = [(x.shape, x.dtype, b is not None) for x, b in zip(xs, bs)]
        ctx.save_for_backward(*saved)
        return y.view(*lead, N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        be = get_backend()
        n = ctx.nterms
        saved = ctx.saved_tensors
        N = saved[1].shape[0]
        dyo = _operand(_rows(dy, N))
        M = dyo.shape[0]
        dxs, dws, dbs = [], [], []
        db_shared = None
        biases = ctx.biases
        if ctx.sm_limit:
            be.set_gemm_sm_limit(ctx.sm_limit)
        for i in range(n):
            xo, w = saved[2 * i], saved[2 * i + 1]
            xshape, xdtype, has_b = ctx.meta[i]
            K = w.shape[1]
            dx = dw = db = None
            if ctx.needs_input_grad[2 + i] and ctx.sinks[i] is not None:
                ctx.sinks[i].accumulate(lambda buf, acc, w=w: be.linear_bwd_data(dyo, _operand(w.detach(), True), buf, accumulate=acc),
                                        (M, K), xdtype, dy.device)
            elif ctx.needs_input_grad[2 + i]:
                dx = torch.empty(M, K, dtype=xdtype, device=dy.device)
                be.linear_bwd_data(dyo, _operand(w.detach(), True), dx)
                dx = dx.view(xshape)
            if ctx.needs_input_grad[2 + n + i]:
                want_b = has_b and ctx.needs_input_grad[2 + 2 * n + i]
                b = biases[i]
                if _fuse_grads and w.grad is not None and (not want_b or b.grad is not None):
                    with _on_leaf(dyo, xo):
                        be.linear_bwd_weight(dyo, xo, w.grad, b.grad if want_b else None, accumulate=True)
                else:
                    dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
                    if want_b and db_shared is None:
                        db_shared = torch.empty(N, dtype=torch.float32, device=dy.device)
                        be.linear_bwd_weight(dyo, xo, dw, db_shared)
                        db = db_shared
                    else:
                        be.linear_bwd_weight(dyo, xo, dw, None)
                        db = db_shared.clone() if want_b else None
            dxs.append(dx)
            dws.append(dw)
            dbs.append(db)
        if ctx.sm_limit:
            be.set_gemm_sm_limit(0)
        return (None, None, *dxs, *dws, *dbs, *([None] * n))


def linear_sum(terms, out_bf16: bool = False):
    """terms: list of (x, weight, bias) or (x, weight, bias, x_op)."""
    xs, ws, bs = zip(*[t[:3] for t in terms])
    x_ops = [t[3] if len(t) > 3 else None for t in terms]
    return LinearSumFn.apply(out_bf16, len(terms), *xs, *ws, *bs, *x_ops)


class LinearGroupFn(Function):
    """Several Linear layers over shared inputs in ONE kernel launch per pass (stcat_linear_group):
        y_j = act_j( sum_{t in job j} x_{i(t)} W_t[r0:r1]^T + b_t[r0:r1] )
    The decoder's query-side layers are chains of [t, 256] GEMMs bounded by launch latency; grouping the independent
    ones (q / k / v projections, query_decoder.py:329-342) and summing the added ones in one accumulator turns ~10
    forward and ~25 backward launches per self-attention block into 2 and 6.  Backward: one grouped dgrad launch
    (dx_i = sum over the terms that read x_i), one grouped wgrad launch, one grouped bias column-sum launch."""

    @staticmethod
    def forward(ctx, spec, *tensors):
        be = get_backend()
        nin, jobs = spec["nin"], spec["jobs"]
        xs, x_ops = tensors[:nin], tensors[nin:2 * nin]
        nt = sum(len(j["terms"]) for j in jobs)
        ws, bs = tensors[2 * nin:2 * nin + nt], tensors[2 * nin + nt:2 * nin + 2 * nt]
        bf = _precision == "bf16"
        xos, M = [], None
        for x, xo in zip(xs, x_ops):
            K = x.shape[-1]
            x2 = _rows(x.detach(), K)
            M = x2.shape[0] if M is None else M
            assert x2.shape[0] == M, "inputs of a grouped Linear share their row count"
            xos.append(_rows(xo.detach(), K) if (xo is not None and bf) else _operand(x2))
        outs, calls, ti = [], {}, 0
        wops = []
        for j in jobs:
            terms = []
            for (i, rows) in j["terms"]:
                w = ws[ti].detach()
                b = None if bs[ti] is None else bs[ti].detach()
                if rows is not None:
                    w = w[rows[0]:rows[1]]
                    b = None if b is None else b[rows[0]:rows[1]]
                wo = _operand(w, True)
                wops.append(wo)
                terms.append((xos[i], wo, b))
                ti += 1
            N = terms[0][1].shape[0]
            odt = torch.bfloat16 if (j.get("out_bf16") and bf) else torch.float32
            y = torch.empty(M, N, dtype=odt, device=xs[0].device)
            outs.append(y)
            calls.setdefault(odt, []).append(dict(terms=terms, out=y, relu=bool(j.get("relu"))))
        ctx.sm_limit = _sm_limit
        with _capped(be, ctx.sm_limit):
            for group in calls.values():
                be.linear_group(0, group)
        ctx.spec = spec
        ctx.biases = bs
        ctx.sinks = [(_sink_of(x) if x.dim() == 2 else None) for x in xs]
        ctx.x_meta = [(x.shape, x.dtype) for x in xs]
        ctx.save_for_backward(*xos, *ws, *[y if j.get("relu") else None for y, j in zip(outs, jobs)])
        lead = xs[0].shape[:-1]
        return tuple(y.view(*lead, y.shape[-1]) for y in outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *dys):
        be = get_backend()
        spec = ctx.spec
        nin, jobs = spec["nin"], spec["jobs"]
        nt = sum(len(j["terms"]) for j in jobs)
        saved = ctx.saved_tensors
        xos, ws, relu_ys = saved[:nin], saved[nin:nin + nt], saved[nin + nt:]
        bs = ctx.biases
        f32 = torch.float32
        dev = xos[0].device
        M = xos[0].shape[0]
        # upstream gradients as GEMM operands (ReLU backward first where the job had one)
        dyos = []
        for jx, (j, dy) in enumerate(zip(jobs, dys)):
            N = dy.shape[-1]
            dy2 = _rows(dy, N)
            if j.get("relu"):
                if not (relu_ys[jx].dtype == torch.bfloat16 and dy2.dtype == torch.bfloat16):  # see LinearFn.backward
                    dy2 = dy2.clone() if dy2.data_ptr() == dy.data_ptr() else dy2
                be.relu_bwd(relu_ys[jx], dy2)
            dyos.append(_operand(dy2))
        # term table: (job index, input index, weight slice view, bias, rows)
        table, ti = [], 0
        for jx, j in enumerate(jobs):
            for (i, rows) in j["terms"]:
                table.append((jx, i, ws[ti], bs[ti], rows))
                ti += 1

        def wslice(w, rows):
            wd = w.detach()
            return wd if rows is None else wd[rows[0]:rows[1]]

        # ---- dgrad: dx_i = sum over terms reading x_i of dy_j . W_t ----
        dxs = [None] * nin
        dcalls = {}
        sinks_used = []
        for i in range(nin):
            if not ctx.needs_input_grad[1 + i]:
                continue
            terms = [(dyos[jx], _operand(wslice(w, rows), True), None) for (jx, ii, w, b, rows) in table if ii == i]
            if not terms:
                continue
            shape, xdt = ctx.x_meta[i]
            sink = ctx.sinks[i]
            if sink is not None:
                # the input is read by many Linear nodes (_GradSink): accumulate into its buffer, hand autograd nothing
                first = sink.buf is None
                if first:
                    sink.buf = torch.empty(M, shape[-1], dtype=xdt, device=dev)
                dx = sink.buf
                sinks_used.append(sink)
            else:
                dx = torch.empty(M, shape[-1], dtype=xdt, device=dev)
                dxs[i] = dx.view(shape)
                first = True
            while terms:  # more than three readers of one input: further launches accumulate
                dcalls.setdefault((xdt, first), []).append(dict(terms=terms[:3], out=dx, accumulate=not first))
                terms = terms[3:]
                first = False
        with _capped(be, ctx.sm_limit):
            for (_, first), group in sorted(dcalls.items(), key=lambda kv: not kv[0][1]):
                be.linear_group(1, group)
        for sink in sinks_used:
            if sink.buf.is_cuda:
                sink.event = torch.cuda.current_stream().record_event()
        # ---- wgrad + bias column sums ----
        dws, dbs = [None] * nt, [None] * nt
        wjobs = []
        all_fused = True
        for tx, (jx, i, w, b, rows) in enumerate(table):
            if not ctx.needs_input_grad[1 + 2 * nin + tx]:
                continue
            want_b = b is not None and ctx.needs_input_grad[1 + 2 * nin + nt + tx]
            if _fuse_grads and w.grad is not None and (not want_b or b.grad is not None):
                gw = w.grad if rows is None else w.grad[rows[0]:rows[1]]
                gb = None if not want_b else (b.grad if rows is None else b.grad[rows[0]:rows[1]])
            else:
                all_fused = False
                dws[tx] = torch.zeros(w.shape, dtype=f32, device=dev)
                gw = dws[tx] if rows is None else dws[tx][rows[0]:rows[1]]
                gb = None
                if want_b:
                    dbs[tx] = torch.zeros(b.shape, dtype=f32, device=dev)
                    gb = dbs[tx] if rows is None else dbs[tx][rows[0]:rows[1]]
            wjobs.append(dict(terms=[(dyos[jx], xos[i], None)], out=gw, accumulate=True, dbias=gb))
        with _on_leaf(*((dyos + [x for x in xos if x is not None]) if all_fused else [])), _capped(be, ctx.sm_limit):
            for k in range(0, len(wjobs), 12):
                be.linear_group(2, wjobs[k:k + 12])
        return (None, *dxs, *([None] * nin), *dws, *dbs)


def linear_group(inputs, jobs):
    """inputs: list of (x, x_op-or-None); jobs: list of dicts {"terms": [(input index, weight, bias, rows-or-None), ...],
    "relu": bool, "out_bf16": bool}.  Returns one output per job."""
    spec = {"nin": len(inputs), "jobs": [{"terms": [(t[0], t[3] if len(t) > 3 else None) for t in j["terms"]],
                                          "relu": j.get("relu", False), "out_bf16": j.get("out_bf16", False)} for j in jobs]}
    xs = [x for x, _ in inputs]
    x_ops = [xo for _, xo in inputs]
    ws = [t[1] for j in jobs for t in j["terms"]]
    bs = [t[2] for j in jobs for t in j["terms"]]
    return LinearGroupFn.apply(spec, *xs, *x_ops, *ws, *bs)


class LayerNormFn(Function):
    """y = LayerNorm(x + res) over the last dim (d = 256), eps 1e-5; fp32 in / out."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps: float, want_op: bool = False):
        be = get_backend()
        d = x.shape[-1]
        x2 = _rows(x.detach().float(), d)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        r2 = None
        if res is not None:
            r2 = _rows(res.detach().float(), d)
            r2 = r2 if r2.is_contiguous() else r2.contiguous()
        rows = x2.shape[0]
        y = torch.empty(rows, d, dtype=torch.float32, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        y_op = torch.empty(rows, d, dtype=torch.bfloat16, device=x.device) if (want_op and _precision == "bf16") else None
        be.layernorm_fwd(x2, r2, gamma.detach(), beta.detach(), y, y_op, mean, rstd, eps)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x2, r2, gamma, mean, rstd)
        ctx.beta = beta
        ctx.shape = x.shape
        ctx.dtypes = (x.dtype, None if res is None else res.dtype)
        if not want_op:
            return y.view(x.shape)
        if y_op is not None:
            y_op = y_op.view(x.shape)
            ctx.mark_non_differentiable(y_op)
        return y.view(x.shape), y_op

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, *_unused):
        if dy is None:
            return (None,) * 6
        be = get_backend()
        x2, r2, gamma, mean, rstd = ctx.saved_tensors
        d = x2.shape[1]
        dy2 = _rows(dy, d)
        dy2 = dy2 if dy2.is_contiguous() else dy2.contiguous()
        dz = torch.empty_like(x2)
        dg, db, _ = _ln_bwd(be, dy2, x2, r2, gamma, ctx.beta, mean, rstd, dz)
        dzv = dz.view(ctx.shape)
        dx = dzv.to(ctx.dtypes[0]) if ctx.needs_input_grad[0] else None
        dr = dzv.to(ctx.dtypes[1]) if (r2 is not None and ctx.needs_input_grad[1]) else None
        return dx, dr, dg, db, None, None


def layer_norm(x, res, gamma, beta, eps: float = 1e-5, want_op: bool = False):
    """LayerNorm(x + res).  With ``want_op`` returns (y, y_op): y_op is the bf16 operand copy written by the same
    kernel (None in fp32 mode) for the Linears that consume y."""
    if want_op:
        return LayerNormFn.apply(x, res, gamma, beta, eps, True)
    return LayerNormFn.apply(x, res, gamma, beta, eps)


class AttentionFn(Function):
    """Batch-major multi-head attention core; see include/stcat_b200.h (stcat_attention_fwd)."""

    @staticmethod
    def forward(ctx, q1, q2, k1, k2, v, key_mask, B, H, Lq, Lk, scale, need_pavg, drop_p=0.0):
        be = get_backend()
        E = H * 32
        ctx.drop = None
        if drop_p:
            seed, off = _drop_reserve(B * H * Lq * Lk)
            ctx.drop = (float(drop_p), seed, off)
        two = q2 is not None

        def prep(t, L):
            t2 = t.detach()
            assert t2.shape == (B * L, E), (t2.shape, B, L, E)
            if _precision == "bf16" and t2.dtype != torch.bfloat16:
                t2 = _operand(t2)
            elif t2.stride(1) != 1:
                t2 = t2.contiguous()
            return t2

        tq1, tk1, tv = prep(q1, Lq), prep(k1, Lk), prep(v, Lk)
        tq2, tk2 = (prep(q2, Lq), prep(k2, Lk)) if two else (None, None)
        if two:  # the kernel takes one leading dimension per operand pair
            if tq2.stride(0) != tq1.stride(0):
                tq1, tq2 = tq1.contiguous(), tq2.contiguous()
            if tk2.stride(0) != tk1.stride(0):
                tk1, tk2 = tk1.contiguous(), tk2.contiguous()
        o = torch.empty(B * Lq, E, dtype=tq1.dtype, device=q1.device)
        lse = torch.empty(B, H, Lq, dtype=torch.float32, device=q1.device)
        pavg = torch.zeros(B, Lq, Lk, dtype=torch.float32, device=q1.device) if need_pavg else None
        be.attention_fwd(tq1, tq2, tk1, tk2, tv, o, key_mask, lse, pavg, B, H, Lq, Lk, scale, drop=ctx.drop)
        ctx.save_for_backward(tq1, tq2, tk1, tk2, tv, key_mask, lse)
        ctx.dims = (B, H, Lq, Lk, scale)
        ctx.in_dtypes = (q1.dtype, k1.dtype, v.dtype)
        if need_pavg:
            return o, pavg
        ctx.mark_non_differentiable()
        return o, None

    @staticmethod
    @once_differentiable
    def backward(ctx, d_o, d_pavg):
        be = get_backend()
        tq1, tq2, tk1, tk2, tv, key_mask, lse = ctx.saved_tensors
        B, H, Lq, Lk, scale = ctx.dims
        E = H * 32
        two = tq2 is not None
        dt = tq1.dtype
        g = d_o.detach()
        if g.dtype != dt:
            g = _operand(g) if dt == torch.bfloat16 else g.float()
        g = _rows(g, E)
        if d_pavg is not None:
            d_pavg = d_pavg.contiguous().float()
        delta = torch.empty(B, H, Lq, dtype=torch.float32, device=g.device)
        mk = lambda L: torch.empty(B * L, E, dtype=dt, device=g.device)
        dq1, dk1, dv = mk(Lq), mk(Lk), mk(Lk)
        dq2, dk2 = (mk(Lq), mk(Lk)) if two else (None, None)
        be.attention_bwd(tq1, tq2, tk1, tk2, tv, g, key_mask, lse, d_pavg, delta, dq1, dq2, dk1, dk2, dv, B, H, Lq, Lk,
                         scale, drop=ctx.drop)
        qd, kd, vd = ctx.in_dtypes
        cast = lambda t, d: None if t is None else (t if t.dtype == d else t.to(d))
        return (cast(dq1, qd), cast(dq2, qd), cast(dk1, kd), cast(dk2, kd), cast(dv, vd), None, None, None, None, None,
                None, None, None)


def attention(q1, k1, v, B, H, Lq, Lk, scale, key_mask=None, q2=None, k2=None, need_pavg=False, drop_p=0.0):
    """Returns (o [B*Lq, H*32], p_avg [B,Lq,Lk] or None).  key_mask: uint8 [B, Lk], nonzero = masked.  drop_p > 0:
    train-mode dropout on the probabilities (the returned p_avg then averages the dropped probabilities)."""
    return AttentionFn.apply(q1, q2, k1, k2, v, key_mask, B, H, Lq, Lk, float(scale), need_pavg, float(drop_p))


class AddFn(Function):
    """out = a + b for same-shape fp32 tensors (q = k = src + pos); with ``as_operand`` the sum is
    written directly in the GEMM-operand dtype (bf16 in bf16 mode) by the same kernel."""

    @staticmethod
    def forward(ctx, a, b, as_operand):
        be = get_backend()
        a2 = a.detach().contiguous()
        b2 = b.detach().contiguous()
        if as_operand and _precision == "bf16":
            out = torch.empty(a2.shape, dtype=torch.bfloat16, device=a2.device)
            be.add(a2, b2, None, out)
        else:
            out = torch.empty_like(a2)
            be.add(a2, b2, out, None)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g if g.dtype == torch.float32 else g.float()
        return g, g, None


def add(a, b, as_operand: bool = False):
    assert a.shape == b.shape and a.dtype == torch.float32 and b.dtype == torch.float32
    return AddFn.apply(a, b, as_operand)


class CastFn(Function):
    """fp32 -> GEMM operand in the active precision, hoisted out of the consumers (one cast kernel for a
    tensor that feeds many GEMMs, e.g. the encoder memory read by all 12 decoder layers).  Backward
    hands the fp32 gradient through unchanged (consumers produce fp32 dx)."""

    @staticmethod
    def forward(ctx, x):
        x2 = x.detach()
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        out = torch.empty(x2.shape, dtype=torch.bfloat16, device=x.device)
        get_backend().cast_bf16(x2.view(-1, x2.shape[-1]), out.view(-1, x2.shape[-1]))
        return out

    @staticmethod
    def backward(ctx, g):
        return g.float() if g.dtype != torch.float32 else g


def to_operand(x):
    """Identity in fp32 mode; one fp32->bf16 cast kernel in bf16 mode."""
    if _precision == "fp32" or x.dtype == torch.bfloat16:
        return x
    return CastFn.apply(x)


def _new(rows, cols, dtype, like):
    return torch.empty(rows, cols, dtype=dtype, device=like.device)


def _opdtype():
    return torch.bfloat16 if _precision == "bf16" else torch.float32


def _cast_op(be, t):
    """fp32 [R,C] contiguous -> operand dtype (no-op in fp32 mode)."""
    if _precision == "fp32":
        return t
    out = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    be.cast_bf16(t, out)
    return out


class SelfAttnBlockFn(Function):
    """y = LayerNorm(x + MHA(q = k = x + pos, v = x)) for B sequences of L tokens, rows batch-major.

    One fused forward/backward pair for the attention half of the post-norm encoder layer
    (reference modal_encoder.py:228-238 via torch nn.MultiheadAttention, functional.py:5798-5873,
    6630-6665): packed in-projection (q,k from x+pos: one N=512 GEMM; v from x), softmax(QK^T/sqrt(dh)
    + key-padding mask) V without materialising P, out-projection, residual, LayerNorm.
    x [B*L, d] fp32; pos [B*L, d] fp32; key_mask uint8 [B, L] or None.
    Returns (y fp32, y_op): y_op is the bf16 GEMM-operand copy of y in bf16 mode, None in fp32 mode.
    """

    @staticmethod
    def forward(ctx, x, x_op, pos, key_mask, w_in, b_in, w_out, b_out, gamma, beta, B, L, H, eps, pos_cls=None, drop_p=0.0, qk_op=None,
                pre=None):
        be = get_backend()
        ctx.has_pos_cls = pos_cls is not None
        # train-mode dropout (modal_encoder.py:212,237): on the attention probabilities and on the block output before
        # the residual; (p, seed, offset) per site, regenerated in backward
        ctx.drop_attn = ctx.drop_out = None
        if drop_p:
            R_ = x.shape[0]
            ctx.drop_attn = (float(drop_p),) + _drop_reserve(B * H * L * L)
            ctx.drop_out = (float(drop_p),) + _drop_reserve(R_ * x.shape[1])
        R, d = x.shape
        assert R == B * L and x.is_contiguous() and pos.shape == x.shape
        od = _opdtype()
        bf = od == torch.bfloat16
        xd, posd = x.detach(), pos.detach()
        posd = posd if posd.is_contiguous() else posd.contiguous()
        bits = side = None
        if ctx.drop_attn:  # forked first: the (ALU-bound) generator runs under the HBM-bound add and the projection GEMMs
            bits, side = _attn_keep_bits(be, B, H, L, ctx.drop_attn, xd)
        use_pre = bf and pre is not None and x_op is not None and x_op.dtype == od
        if use_pre:
            # q/k/v of every row but row 0 of each sequence were projected EARLY (qkv_early: on a side stream, under the temporal
            # layer that only rewrites those rows); cls_scatter has patched row 0 of qk_in and of x_op: project those B rows now
            qk_in, qkv_pre = pre["qk_in"], pre["qkv"]
            xo = x_op.detach()
            if pre["done"] is not None:
                torch.cuda.current_stream().wait_event(pre["done"])
        elif bf:
            if qk_op is not None and qk_op.dtype == od:  # the producer of x already wrote bf16(x + pos) (ops.token_assembly)
                qk_in = qk_op.detach().view(R, d)
            else:
                qk_in = _new(R, d, od, x)
                be.add(xd, posd, None, qk_in)
            xo = x_op.detach() if (x_op is not None and x_op.dtype == od) else _cast_op(be, xd)
        else:
            qk_in = _new(R, d, od, x)
            be.add(xd, posd, qk_in, None)
            xo = xd
        wi = _operand(w_in.detach(), True)
        wo = _operand(w_out.detach(), True)
        bi = b_in.detach()
        if use_pre:
            qkv = qkv_pre
            q3, x3, o3 = qk_in.view(B, L, d)[:, 0, :], xo.view(B, L, d)[:, 0, :], qkv.view(B, L, 3 * d)[:, 0, :]
            be.linear_group(0, [dict(terms=[(q3, wi[: 2 * d], bi[: 2 * d])], out=o3[:, : 2 * d]),
                                dict(terms=[(x3, wi[2 * d:], bi[2 * d:])], out=o3[:, 2 * d:])])
        else:
            qkv = _new(R, 3 * d, od, x)
            # q, k from x + pos (one N = 512 GEMM) and v from x: two jobs of one grouped launch
            be.linear_group(0, [dict(terms=[(qk_in, wi[: 2 * d], bi[: 2 * d])], out=qkv[:, : 2 * d]),
                                dict(terms=[(xo, wi[2 * d:], bi[2 * d:])], out=qkv[:, 2 * d:])])
        o = _new(R, d, od, x)
        lse = torch.empty(B, H, L, dtype=torch.float32, device=x.device)
        scale = float(d // H) ** -0.5
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        if bits is not None:
            be.attention_fwd(qkv[:, :d], None, qkv[:, d:2 * d], None, qkv[:, 2 * d:], o, key_mask, lse, None, B, H, L, L,
                             scale, drop=ctx.drop_attn, bits=bits)
        else:
            be.attention_fwd(qkv[:, :d], None, qkv[:, d:2 * d], None, qkv[:, 2 * d:], o, key_mask, lse, None, B, H, L, L,
                             scale, drop=ctx.drop_attn)
        ctx.keep_bits = bits
        a = _new(R, d, torch.float32, x)
        be.linear_fwd(o, wo, b_out.detach(), a)
        # the block-output dropout rides in the LayerNorm kernels in bf16 mode (mask on load forward; masked operand copy and
        # bias column sums backward): no separate pass over a / dz.  ``a`` is saved UNdropped in that case.
        ctx.ln_drop = ctx.drop_out if bf else None
        if ctx.drop_out and not bf:
            be.dropout(a, a, *ctx.drop_out)
        y = _new(R, d, torch.float32, x)
        y_op = _new(R, d, od, x) if bf else None
        mean = torch.empty(R, dtype=torch.float32, device=x.device)
        rstd = torch.empty(R, dtype=torch.float32, device=x.device)
        be.layernorm_fwd(a, xd, gamma.detach(), beta.detach(), y, y_op, mean, rstd, eps, drop=ctx.ln_drop)
        ctx.set_materialize_grads(False)  # no full-size zero tensor for the non-differentiable y_op output
        ctx.save_for_backward(xd, xo, qk_in, qkv, o, lse, a, mean, rstd, key_mask, w_in, w_out, gamma)
        ctx.extra = (b_in, b_out, beta)
        ctx.pos_cls = pos_cls
        ctx.dims = (B, L, H, scale)
        if bf:
            ctx.mark_non_differentiable(y_op)
        return y, y_op

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _unused):
        if dy is None:
            return (None,) * 18
        be = get_backend()
        xd, xo, qk_in, qkv, o, lse, a, mean, rstd, key_mask, w_in, w_out, gamma = ctx.saved_tensors
        B, L, H, scale = ctx.dims
        R, d = xd.shape
        od = qkv.dtype
        f32 = torch.float32
        dy = dy if (dy.is_contiguous() and dy.dtype == f32) else dy.contiguous().float()
        wi = _operand(w_in.detach(), True)
        wo = _operand(w_out.detach(), True)
        b_in, b_out, beta = ctx.extra
        dz = _new(R, d, f32, dy)  # grad wrt (a) and wrt the residual x
        bf = od == torch.bfloat16
        if ctx.drop_out and not ctx.ln_drop:
            # the out-projection branch sees dz through the dropout mask of the forward; the residual branch sees dz itself
            dg, dbt, _ = _ln_bwd(be, dy, a, xd, gamma, beta, mean, rstd, dz)
            dza = _new(R, d, f32, dy)
            be.dropout(dz, dza, *ctx.drop_out)
            dz_op = _cast_op(be, dza)
            dwo, dbo = _wgrad(be, dz_op, o, w_out, b_out, True, True)
        else:
            # LayerNorm backward writes the bf16 operand copy itself (masked, with fused dropout: dz_op = mask(dz))
            dz_op = _new(R, d, od, dy) if bf else dz
            dg, dbt, dbo = _ln_bwd(be, dy, a, xd, gamma, beta, mean, rstd, dz, dz_op if bf else None, b_out, drop=ctx.ln_drop)
            dwo, _ = _wgrad(be, dz_op, o, w_out, None, True, False)
        d_o = _new(R, d, od, dy)
        be.linear_bwd_data(dz_op, wo, d_o)
        dqkv = _new(R, 3 * d, od, dy)
        delta = torch.empty(B, H, L, dtype=f32, device=dy.device)
        if ctx.keep_bits is not None:
            be.attention_bwd(qkv[:, :d], None, qkv[:, d:2 * d], None, qkv[:, 2 * d:], d_o, key_mask, lse, None, delta,
                             dqkv[:, :d], None, dqkv[:, d:2 * d], None, dqkv[:, 2 * d:], B, H, L, L, scale, o=o,
                             drop=ctx.drop_attn, bits=ctx.keep_bits)
        else:
            be.attention_bwd(qkv[:, :d], None, qkv[:, d:2 * d], None, qkv[:, 2 * d:], d_o, key_mask, lse, None, delta,
                             dqkv[:, :d], None, dqkv[:, d:2 * d], None, dqkv[:, 2 * d:], B, H, L, L, scale, o=o,
                             drop=ctx.drop_attn)
        if _fuse_grads and w_in.grad is not None and b_in.grad is not None:
            gwi, gbi = w_in.grad, b_in.grad
            dwi = dbi = None
        else:
            gwi = dwi = torch.zeros(3 * d, d, dtype=f32, device=dy.device)
            gbi = dbi = torch.zeros(3 * d, dtype=f32, device=dy.device)
        with _on_leaf(*([dqkv, qk_in, xo] if dwi is None else [])):
            be.linear_group(2, [dict(terms=[(dqkv[:, : 2 * d], qk_in, None)], out=gwi[: 2 * d], accumulate=True, dbias=gbi[: 2 * d]),
                                dict(terms=[(dqkv[:, 2 * d:], xo, None)], out=gwi[2 * d:], accumulate=True, dbias=gbi[2 * d:])])
        need_pos = ctx.needs_input_grad[2]
        dpos = None
        if need_pos:
            dpos = _new(R, d, f32, dy)
            be.linear_bwd_data(dqkv[:, : 2 * d], wi[: 2 * d], dpos)
            be.add(dz, dpos, dz, None)
            be.linear_bwd_data(dqkv[:, 2 * d:], wi[2 * d:], dz, accumulate=True)
        else:
            # dz += dqk Wqk + dv Wv: one two-term accumulating job
            be.linear_group(1, [dict(terms=[(dqkv[:, : 2 * d], wi[: 2 * d], None), (dqkv[:, 2 * d:], wi[2 * d:], None)],
                                     out=dz, accumulate=True)])
        dpc = None
        if ctx.has_pos_cls and ctx.needs_input_grad[14]:
            # pos is constant except its row 0 of every sequence, which is the learned token `pos_cls` ([1, d]):
            # d pos_cls = (sum over sequences of dq,dk at row 0) . Wqk -- a [1, 512] x [512, 256] product instead of a
            # full-size dpos GEMM, add and gradient accumulation per layer.  It is a leaf gradient: with the flat gradient
            # buffer it is accumulated in place on the leaf stream, off the dependent chain of the backward pass.
            pc = ctx.pos_cls
            fused = _fuse_grads and pc is not None and pc.grad is not None and pc.grad.is_contiguous()
            with _on_leaf(*([dqkv] if fused else [])):
                srow = dqkv.view(B, L, 3 * d)[:, 0, : 2 * d].float().sum(0, keepdim=True)
                if fused:
                    be.linear_bwd_data(srow, w_in.detach()[: 2 * d], pc.grad.view(1, d), accumulate=True)
                else:
                    dpc = torch.empty(1, d, dtype=f32, device=dy.device)
                    be.linear_bwd_data(srow, w_in.detach()[: 2 * d], dpc)
        return dz, None, dpos, None, dwi, dbi, dwo, dbo, dg, dbt, None, None, None, None, dpc, None, None, None


class OutLNFn(Function):
    """y = LayerNorm(res + drop(o W^T + b)): out-projection of an attention block + block-output dropout + residual + norm
    (query_decoder.py:342-345, 430-432, 610-613, 652-654) as one autograd node, like the tail of SelfAttnBlockFn: the
    LayerNorm backward hands the (masked) bf16 operand copy of dz straight to the out-projection's gradient GEMMs and sums its
    bias gradient -- no cast, no column-sum and, in train mode, no stand-alone dropout launches on the decoders' chains.
    o [R, d] (operand dtype or fp32), res [R, d] fp32.  Returns (y fp32, y_op)."""

    @staticmethod
    def forward(ctx, o, res, w, b, gamma, beta, eps, drop_p=0.0):
        be = get_backend()
        R, d = res.shape
        od = _opdtype()
        bf = od == torch.bfloat16
        o2 = o.detach()
        o2 = o2 if o2.is_contiguous() else o2.contiguous()
        oo = o2 if o2.dtype == od else _operand(o2)
        resd = res.detach()
        resd = resd if (resd.is_contiguous() and resd.dtype == torch.float32) else resd.contiguous().float()
        wo = _operand(w.detach(), True)
        a = _new(R, d, torch.float32, resd)
        be.linear_fwd(oo, wo, b.detach(), a)
        ctx.drop = ((float(drop_p),) + _drop_reserve(R * d)) if drop_p else None
        ctx.ln_drop = ctx.drop if bf else None
        if ctx.drop and not bf:
            be.dropout(a, a, *ctx.drop)
        y = _new(R, d, torch.float32, resd)
        y_op = _new(R, d, od, resd) if bf else None
        mean = torch.empty(R, dtype=torch.float32, device=resd.device)
        rstd = torch.empty(R, dtype=torch.float32, device=resd.device)
        be.layernorm_fwd(a, resd, gamma.detach(), beta.detach(), y, y_op, mean, rstd, eps, drop=ctx.ln_drop)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(oo, a, resd, mean, rstd, w, gamma)
        ctx.extra = (b, beta)
        ctx.o_dtype = o.dtype
        if bf:
            ctx.mark_non_differentiable(y_op)
        return y, y_op

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _unused):
        if dy is None:
            return (None,) * 8
        be = get_backend()
        oo, a, resd, mean, rstd, w, gamma = ctx.saved_tensors
        b, beta = ctx.extra
        R, d = resd.shape
        f32 = torch.float32
        bf = oo.dtype == torch.bfloat16
        dy = dy if (dy.is_contiguous() and dy.dtype == f32) else dy.contiguous().float()
        wo = _operand(w.detach(), True)
        dz = _new(R, d, f32, dy)
        if ctx.drop and not ctx.ln_drop:
            dg, dbt, _ = _ln_bwd(be, dy, a, resd, gamma, beta, mean, rstd, dz)
            dza = _new(R, d, f32, dy)
            be.dropout(dz, dza, *ctx.drop)
            dz_op = _cast_op(be, dza)
            dw, db = _wgrad(be, dz_op, oo, w, b, True, True)
        else:
            dz_op = _new(R, d, oo.dtype, dy) if bf else dz
            dg, dbt, db = _ln_bwd(be, dy, a, resd, gamma, beta, mean, rstd, dz, dz_op if bf else None, b, drop=ctx.ln_drop)
            dw, _ = _wgrad(be, dz_op, oo, w, None, True, False)
        d_o = None
        if ctx.needs_input_grad[0]:
            d_o = _new(R, d, ctx.o_dtype, dy)
            be.linear_bwd_data(dz_op, wo, d_o)
        return d_o, dz, dw, db, dg, dbt, None, None


def out_ln(o, res, w, b, gamma, beta, eps: float = 1e-5, drop_p: float = 0.0):
    """(y, y_op) = LayerNorm(res + dropout(o W^T + b)) -- see OutLNFn."""
    return OutLNFn.apply(o, res, w, b, gamma, beta, eps, float(drop_p))


class FFNBlockFn(Function):
    """y = LayerNorm(x + W2 relu(W1 x + b1) + b2)  (modal_encoder.py:239-241; query_decoder.py:435-437,
    657-659).  x [R, d] fp32.  Returns (y, y_op) like SelfAttnBlockFn."""

    @staticmethod
    def forward(ctx, x, x_op, w1, b1, w2, b2, gamma, beta, eps, drop_p=0.0):
        be = get_backend()
        R, d = x.shape
        F_ = w1.shape[0]
        # train-mode dropout (modal_encoder.py:239-240; query_decoder.py:435-436, 657-658): on relu(linear1) and on the
        # block output before the residual
        ctx.drop_h = ctx.drop_out = None
        if drop_p:
            ctx.drop_h = (float(drop_p),) + _drop_reserve(R * F_)
            ctx.drop_out = (float(drop_p),) + _drop_reserve(R * d)
        od = _opdtype()
        bf = od == torch.bfloat16
        xd = x.detach()
        xd = xd if xd.is_contiguous() else xd.contiguous()
        if bf:
            xo = x_op.detach() if (x_op is not None and x_op.dtype == od) else _cast_op(be, xd)
        else:
            xo = xd
        w1o, w2o = _operand(w1.detach(), True), _operand(w2.detach(), True)
        h = _new(R, F_, od, x)
        if ctx.drop_h and bf:
            be.linear_dropout_fwd(xo, w1o, b1.detach(), h, True, ctx.drop_h)  # the inner dropout's mask is drawn in the GEMM epilogue
        else:
            be.linear_fwd(xo, w1o, b1.detach(), h, relu=True)
            if ctx.drop_h:
                be.dropout(h, h, *ctx.drop_h)
        # (few rows, long contraction -- decoder / temporal FFN, [t, 2048] x [2048, 256]: the GEMM entry point splits the
        # contraction over a thread-block cluster and sums the partial tiles in rank order, gemm_tcgen05.cu CLK)
        yl = _new(R, d, torch.float32, x)
        be.linear_fwd(h, w2o, b2.detach(), yl)
        ctx.ln_drop = ctx.drop_out if bf else None  # bf16 mode: the output dropout rides in the LayerNorm kernels (yl saved undropped)
        if ctx.drop_out and not bf:
            be.dropout(yl, yl, *ctx.drop_out)
        y = _new(R, d, torch.float32, x)
        y_op = _new(R, d, od, x) if bf else None
        mean = torch.empty(R, dtype=torch.float32, device=x.device)
        rstd = torch.empty(R, dtype=torch.float32, device=x.device)
        be.layernorm_fwd(yl, xd, gamma.detach(), beta.detach(), y, y_op, mean, rstd, eps, drop=ctx.ln_drop)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(xd, xo, h, yl, mean, rstd, w1, w2, gamma)
        ctx.extra = (b1, b2, beta)
        if bf:
            ctx.mark_non_differentiable(y_op)
        return y, y_op

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _unused):
        if dy is None:
            return (None,) * 10
        be = get_backend()
        xd, xo, h, yl, mean, rstd, w1, w2, gamma = ctx.saved_tensors
        R, d = xd.shape
        F_ = w1.shape[0]
        od = h.dtype
        f32 = torch.float32
        dy = dy if (dy.is_contiguous() and dy.dtype == f32) else dy.contiguous().float()
        w1o, w2o = _operand(w1.detach(), True), _operand(w2.detach(), True)
        b1, b2, beta = ctx.extra
        dz = _new(R, d, f32, dy)
        bf = od == torch.bfloat16
        dh = _new(R, F_, od, dy)
        alpha = 1.0
        if ctx.drop_out and not ctx.ln_drop:
            # dropout in exact-fp32 mode, unfused: the linear2 branch sees dz through the output mask; dh = (dza W2) * relu
            # mask, then the hidden mask; bias gradients from the masked tensors
            dg, dbt, _ = _ln_bwd(be, dy, yl, xd, gamma, beta, mean, rstd, dz)
            dza = _new(R, d, f32, dy)
            be.dropout(dz, dza, *ctx.drop_out)
            dz_op = _cast_op(be, dza)
            dw2, db2 = _wgrad(be, dz_op, h, w2, b2, True, True)
            be.linear_bwd_data(dz_op, w2o, dh, relu_y=h)
            be.dropout(dh, dh, *ctx.drop_h)
            dw1, db1 = _wgrad(be, dh, xo, w1, b1, True, True)
            be.linear_bwd_data(dh, w1o, dz, accumulate=True)
            return dz, None, dw1, db1, dw2, db2, dg, dbt, None, None
        if ctx.drop_out:
            # dropout in bf16 mode rides in the kernels of the dropout-free path: the LayerNorm backward applies the output mask
            # (dz_op = bf16(mask(dz)), db2 fused); the saved h is the DROPPED activation, so the dgrad epilogue's ReLU mask is
            # already the product of both masks and only the 1 / keep factor is left (alpha): no mask is regenerated
            alpha = _drop_scale(ctx.drop_h[0])
        dz_op = _new(R, d, od, dy) if bf else dz
        dg, dbt, db2 = _ln_bwd(be, dy, yl, xd, gamma, beta, mean, rstd, dz, dz_op if bf else None, b2, drop=ctx.ln_drop)
        dw2, _ = _wgrad(be, dz_op, h, w2, None, True, False)
        # dh = (dz W2) * (h > 0) with db1 = colsum(dh): ReLU backward and the bias gradient ride in the dgrad epilogue
        if _fuse_grads and b1.grad is not None:
            db1 = None
            be.linear_bwd_data(dz_op, w2o, dh, relu_y=h, dbias=b1.grad, alpha=alpha)
        else:
            db1 = torch.zeros(F_, dtype=f32, device=dy.device)
            be.linear_bwd_data(dz_op, w2o, dh, relu_y=h, dbias=db1, alpha=alpha)
        dw1, _ = _wgrad(be, dh, xo, w1, None, True, False)
        be.linear_bwd_data(dh, w1o, dz, accumulate=True)
        return dz, None, dw1, db1, dw2, db2, dg, dbt, None, None


def self_attn_block(x, x_op, pos, key_mask, w_in, b_in, w_out, b_out, gamma, beta, B, L, H, eps=1e-5, pos_cls=None, drop_p=0.0,
                    qk_op=None, pre=None):
    """``pos_cls`` ([1, d], optional): the parameter that row 0 of every sequence of ``pos`` was copied from; when
    given, ``pos`` itself is treated as a constant and the gradient goes to ``pos_cls`` directly.  ``qk_op`` (optional): the
    bf16 operand copy of ``x + pos`` when the producer of ``x`` already wrote it."""
    return SelfAttnBlockFn.apply(x, x_op, pos, key_mask, w_in, b_in, w_out, b_out, gamma, beta, B, L, H, eps, pos_cls, float(drop_p),
                                 qk_op, pre)


_early_streams = {}


@torch.no_grad()
def qkv_early(x, x_op, pos, w_in, b_in):
    """The packed q/k/v in-projection of a spatial encoder layer, issued on a side stream BEFORE the temporal layer in front of
    it has run (bf16 mode; x [R, d] fp32, x_op its bf16 copy, pos [R, d]).  The temporal layer only rewrites row 0 of every
    sequence (the frame-CLS rows, modal_encoder.py:191-195), so every other row of bf16(x + pos) and of q/k/v is final; the
    caller hands ``pre["qk_in"]`` to ``cls_scatter`` (which patches its CLS rows after ``pre["added"]``) and ``pre`` to
    ``self_attn_block``, which projects the B patched rows and waits for ``pre["done"]``: the 20 us add + GEMM of every layer
    hide under the latency-bound temporal chain.  Not an autograd node: the block function saves these tensors itself."""
    be = get_backend()
    R, d = x.shape
    xd, xo, posd = x.detach(), x_op.detach(), pos.detach()
    cuda = xd.is_cuda
    if cuda:
        cur = torch.cuda.current_stream()
        side = _early_streams.get(xd.device)
        if side is None:
            side = _early_streams[xd.device] = torch.cuda.Stream(xd.device)
        side.wait_stream(cur)
        ctxm = torch.cuda.stream(side)
    else:
        import contextlib

        ctxm = contextlib.nullcontext()
    added = done = None
    with ctxm:
        qk_in = torch.empty(R, d, dtype=torch.bfloat16, device=xd.device)
        be.add(xd, posd, None, qk_in)
        if cuda:
            added = side.record_event()
        wi, bi = _operand(w_in.detach(), True), b_in.detach()
        qkv = torch.empty(R, 3 * d, dtype=torch.bfloat16, device=xd.device)
        be.linear_group(0, [dict(terms=[(qk_in, wi[: 2 * d], bi[: 2 * d])], out=qkv[:, : 2 * d]),
                            dict(terms=[(xo, wi[2 * d:], bi[2 * d:])], out=qkv[:, 2 * d:])])
        if cuda:
            done = side.record_event()
    if cuda:
        for t in (qk_in, qkv):
            t.record_stream(cur)
        for t in (xd, xo, posd):
            t.record_stream(side)
    return {"qk_in": qk_in, "qkv": qkv, "added": added, "done": done}


def ffn_block(x, x_op, w1, b1, w2, b2, gamma, beta, eps=1e-5, drop_p=0.0):
    return FFNBlockFn.apply(x, x_op, w1, b1, w2, b2, gamma, beta, eps, float(drop_p))


class TakeRowsFn(Function):
    """(base, rows) = (x, x[:, r, :]) for x [n, S, d]: hands out row r of every sequence (the frame-CLS tokens the
    temporal encoder layer works on, modal_encoder.py:170-175) together with x itself.  One node with two outputs so
    that the backward is a row-sized in-place add (d base[:, r, :] += d rows) instead of autograd's full-size zero
    tensor + full-size sum for a select."""

    @staticmethod
    def forward(ctx, x, r: int):
        ctx.r = r
        ctx.set_materialize_grads(False)
        # the base output aliases x's storage (detached alias, not an autograd view: it may be written in place later)
        return x.detach(), x.detach()[:, r, :].clone()

    @staticmethod
    def backward(ctx, g_base, g_rows):
        if g_base is None:
            if g_rows is None:
                return None, None
            raise RuntimeError("TakeRowsFn: the base output must be used (it carries the sequence's gradient)")
        if g_rows is not None:
            g_base = g_base if g_base.is_contiguous() else g_base.contiguous()
            g_base[:, ctx.r, :] += g_rows  # g_base is ours: produced for this node alone by PutRowsFn / the next block
        return g_base, None


class PutRowsFn(Function):
    """x[:, r, :] = rows, in place on x's storage (the reference's ``output[0, :, :] = frames_src``,
    modal_encoder.py:191-195).  Backward: d rows = g[:, r, :]; d x = g with those rows zeroed -- done in place on g
    (row-sized kernels) where autograd's CopySlices clones the full tensor."""

    @staticmethod
    def forward(ctx, x, rows, r: int):
        ctx.r = r
        out = x.detach()  # same storage and version counter as x; a new autograd identity owned by this node
        out[:, r, :] = rows.detach()
        return out

    @staticmethod
    def backward(ctx, g):
        g = g if g.is_contiguous() else g.contiguous()
        g_rows = g[:, ctx.r, :].clone()
        g[:, ctx.r, :] = 0  # g is the gradient tensor produced for this node by the next block's backward: ours to edit
        return g, g_rows, None


class ClsGatherFn(Function):
    """(x alias, Y) with Y = [video ; x[:, r, :]]: ``take_rows`` + ``torch.cat`` of the temporal layer's input for ONE
    un-padded video (x [n, S, d], video [1, d] -> Y [1 + n, d]; modal_encoder.py:170-177).  With ``pos`` ([1 + n, d], bf16 mode)
    one kernel also writes the temporal layer's two GEMM operands bf16(Y + pos) and bf16(Y) (returned as constants).
    Backward: the row gradient is added onto the stream's gradient in place (row-sized), the video token's gradient is a view."""

    @staticmethod
    def forward(ctx, x, video, r: int, pos=None):
        ctx.r = r
        ctx.set_materialize_grads(False)
        xd = x.detach()
        if pos is not None and _precision == "bf16" and xd.is_contiguous() and xd.dtype == torch.float32:
            n, S, d = xd.shape
            Y = torch.empty(1 + n, d, dtype=torch.float32, device=xd.device)
            qk_op = torch.empty(1 + n, d, dtype=torch.bfloat16, device=xd.device)
            y_op = torch.empty(1 + n, d, dtype=torch.bfloat16, device=xd.device)
            get_backend().cls_gather(xd, video.detach().float().contiguous(), pos.detach().contiguous(), Y, qk_op, y_op, r)
            ctx.mark_non_differentiable(qk_op, y_op)
            return xd, Y, qk_op, y_op
        return xd, torch.cat([video.detach(), xd[:, r, :]], 0), None, None

    @staticmethod
    def backward(ctx, g_base, g_y, *_unused):
        if g_base is None:
            if g_y is None:
                return None, None, None, None
            raise RuntimeError("ClsGatherFn: the base output must be used (it carries the sequence's gradient)")
        if g_y is None:
            return g_base, None, None, None
        g_base = g_base if g_base.is_contiguous() else g_base.contiguous()
        g_base[:, ctx.r, :] += g_y[1:]  # g_base is ours: produced for this node alone by ClsScatterFn / the next block
        return g_base, g_y[:1], None, None


class ClsScatterFn(Function):
    """x[:, r, :] = Y[1:] in place on x's storage, video' = Y[:1] (``put_rows`` + the two slices of the temporal layer's
    output, modal_encoder.py:191-195); with ``x_op`` (the stream's bf16 operand copy, bf16 mode) the same kernel refreshes
    its rows too.  Backward: one concatenation builds dY, the replaced rows of dx are zeroed in place."""

    @staticmethod
    def forward(ctx, x, y, r: int, x_op=None, pre=None, pos=None):
        ctx.r = r
        ctx.set_materialize_grads(False)
        out, yd = x.detach(), y.detach()
        if x_op is not None and _precision == "bf16" and out.is_contiguous() and yd.is_contiguous() and x_op.is_contiguous():
            qk_next = None
            if pre is not None:  # ops.qkv_early: its bf16(x + pos) holds the stale rows r; patch them once its add has run
                qk_next = pre["qk_in"].view(out.shape)
                if pre["added"] is not None:
                    torch.cuda.current_stream().wait_event(pre["added"])
            get_backend().cls_scatter(yd, out, x_op.detach().view(out.shape), r, qk_next,
                                      None if qk_next is None else pos.detach().view(out.shape))
        else:
            assert pre is None
            out[:, r, :] = yd[1:]
            if x_op is not None:
                x_op.detach().view(out.shape)[:, r, :] = yd[1:]
        return out, yd[:1]

    @staticmethod
    def backward(ctx, g, g_video):
        if g is None:
            raise RuntimeError("ClsScatterFn: the sequence output must be used")
        g = g if g.is_contiguous() else g.contiguous()
        gv = g_video if g_video is not None else g.new_zeros(1, g.shape[2])
        g_y = torch.cat([gv, g[:, ctx.r, :]], 0)
        g[:, ctx.r, :] = 0  # g is the gradient tensor produced for this node by the next block's backward: ours to edit
        return g, g_y, None, None, None, None


def cls_gather(x, video, r: int = 0, pos=None):
    """(x alias, Y, bf16(Y + pos) or None, bf16(Y) or None)"""
    return ClsGatherFn.apply(x, video, r, pos)


def cls_scatter(x, y, r: int = 0, x_op=None, pre=None, pos=None):
    """(x with rows r replaced by Y[1:], Y[:1]); ``x_op``: the bf16 operand copy of x, refreshed in the same launch; ``pre`` /
    ``pos``: the next layer's early projection (ops.qkv_early), whose bf16(x + pos) rows r are patched in the same launch"""
    return ClsScatterFn.apply(x, y, r, x_op, pre, pos)


def take_rows(x, r: int = 0):
    return TakeRowsFn.apply(x, r)


def put_rows(x, rows, r: int = 0):
    return PutRowsFn.apply(x, rows, r)


class TokenAssemblyFn(Function):
    """Encoder input assembly (modal_encoder.py:40-72) as one launch each way: X[f] = [frame_cls ; vis[f]^T ; text[:, v(f)]],
    POS[f] = [local_pos ; vis_pos[f]^T ; 0] in the frame-major layout [n, S, d], plus the first spatial layer's two GEMM-operand
    copies (bf16(X + POS), bf16(X)).  POS and the operand copies are non-differentiable: the caller routes the gradient of
    ``local_pos`` through ``self_attn_block(pos_cls=...)`` and uses this node only when ``vis_pos`` needs no gradient.
    Backward: d vis (transposed back), d text (sum over the frames of each video), d frame_cls (sum over all frames)."""

    @staticmethod
    def forward(ctx, vis, vis_pos, text, frame_cls, local_pos, f2v, vid_start):
        be = get_backend()
        n, d = vis.shape[0], vis.shape[1]
        HW = vis[0, 0].numel()
        L, b = text.shape[0], text.shape[1]
        S = 1 + HW + L
        dev = vis.device
        X = torch.empty(n, S, d, dtype=torch.float32, device=dev)
        POS = torch.empty(n, S, d, dtype=torch.float32, device=dev)
        qk_op = torch.empty(n, S, d, dtype=torch.bfloat16, device=dev)
        x_op = torch.empty(n, S, d, dtype=torch.bfloat16, device=dev)
        be.token_assembly(vis.detach().contiguous(), vis_pos.detach().contiguous(), text.detach().contiguous(), f2v,
                          frame_cls.detach().contiguous(), local_pos.detach().contiguous(), X, POS, qk_op, x_op)
        ctx.dims = (tuple(vis.shape), tuple(text.shape), tuple(frame_cls.shape), HW, L, b)
        ctx.vid_start = vid_start
        ctx.mark_non_differentiable(POS, qk_op, x_op)
        ctx.set_materialize_grads(False)
        return X, POS, qk_op, x_op

    @staticmethod
    @once_differentiable
    def backward(ctx, gX, *_unused):
        if gX is None:
            return (None,) * 7
        vis_shape, text_shape, cls_shape, HW, L, b = ctx.dims
        gX = gX if (gX.is_contiguous() and gX.dtype == torch.float32) else gX.contiguous().float()
        need_vis, _, need_text, need_cls = ctx.needs_input_grad[:4]
        dev = gX.device
        dvis = torch.empty(vis_shape, dtype=torch.float32, device=dev) if need_vis else None
        dtext = torch.empty(text_shape, dtype=torch.float32, device=dev) if (need_text and L > 0) else None
        dcls = torch.empty(cls_shape, dtype=torch.float32, device=dev) if need_cls else None
        if dvis is not None or dtext is not None or dcls is not None:
            get_backend().token_assembly_bwd(gX, dvis, dtext, dcls, ctx.vid_start, HW, L, b)
        if need_text and dtext is None:
            dtext = torch.zeros(text_shape, dtype=torch.float32, device=dev)
        return dvis, None, dtext, dcls, None, None, None


def token_assembly(vis, vis_pos, text, frame_cls, local_pos, f2v=None, vid_start=None):
    return TokenAssemblyFn.apply(vis, vis_pos, text, frame_cls, local_pos, f2v, vid_start)


class MemOperandsFn(Function):
    """The decoder's views of the encoder stream X [n, S, d] (query_decoder.py:83-96, 355-366, 633-639) in one launch:
    (bf16(X[:, 1:]), bf16(POS[:, 1:]), bf16(X[:, 1:] + POS[:, 1:])) as [n (S-1), d] GEMM operands and the fp32 frame-CLS rows
    X[:, 0].  Backward: dX = [g_cls ; g_mem + g_mempos] in one launch (the reference layout's slice / select / transpose
    nodes cost two full-size zero fills, two copies, two casts and two adds).  POS is a constant here."""

    @staticmethod
    def forward(ctx, X, POS, sinks=None):
        be = get_backend()
        ctx.sinks = sinks
        n, S, d = X.shape
        M = S - 1
        dev = X.device
        bf = torch.bfloat16
        mem_op = torch.empty(n * M, d, dtype=bf, device=dev)
        pos_op = torch.empty(n * M, d, dtype=bf, device=dev)
        mempos_op = torch.empty(n * M, d, dtype=bf, device=dev)
        cls = torch.empty(n, d, dtype=torch.float32, device=dev)
        Xd = X.detach()
        mark_phase("encoder_fwd_end")
        be.mem_operands(Xd if Xd.is_contiguous() else Xd.contiguous(), POS.detach().contiguous(), mem_op, pos_op, mempos_op, cls)
        ctx.shape = (n, S, d)
        ctx.mark_non_differentiable(pos_op)
        ctx.set_materialize_grads(False)
        return mem_op, pos_op, mempos_op, cls

    @staticmethod
    @once_differentiable
    def backward(ctx, g_mem, _g_pos, g_mempos, g_cls):
        if ctx.sinks is not None:  # contributions the consumers accumulated in their GEMMs (_GradSink)
            s_mem, s_mempos = ctx.sinks[0].take(), ctx.sinks[1].take()
            g_mem = s_mem if g_mem is None else (g_mem if s_mem is None else g_mem.float() + s_mem.float())
            g_mempos = s_mempos if g_mempos is None else (g_mempos if s_mempos is None else g_mempos.float() + s_mempos.float())
        if g_mem is None and g_mempos is None and g_cls is None:
            return None, None, None
        n, S, d = ctx.shape
        ref = g_mem if g_mem is not None else (g_mempos if g_mempos is not None else g_cls)
        fix = lambda g: None if g is None else (g if g.is_contiguous() else g.contiguous())
        g_cls = None if g_cls is None else fix(g_cls.float() if g_cls.dtype != torch.float32 else g_cls)
        dX = torch.empty(n, S, d, dtype=torch.float32, device=ref.device)
        get_backend().mem_operands_bwd(fix(g_mem), fix(g_mempos), g_cls, dX)
        mark_phase("decoder_bwd_end")
        return dX, None, None


def mem_operands(X, POS):
    """(mem_op, pos_op, mempos_op, cls); mem_op / mempos_op carry a gradient sink (_GradSink) for their Linear consumers."""
    sinks = (_GradSink(), _GradSink()) if _GRAD_SINK else None
    out = MemOperandsFn.apply(X, POS, sinks)
    if sinks is not None:
        out[0]._stcat_sink, out[2]._stcat_sink = sinks
    return out


class TemplateFn(Function):
    """TemplateGenerator.forward (query_decoder.py:441-475) + the sigmoid of :105: (anchor [n, q], temp_query [n, d]) from the
    video tokens [b, d] and the frame-CLS tokens [n, d]; two launches forward, three backward (bf16 mode)."""

    @staticmethod
    def forward(ctx, videos_cls, frames_cls, Wc, bc, Wg, bg, Wb, bb, Wa, ba, f2v, vid_start):
        be = get_backend()
        n, d = frames_cls.shape
        b, q = videos_cls.shape[0], Wa.shape[0]
        dev = frames_cls.device
        f32 = torch.float32
        v = videos_cls.detach().float().contiguous()
        fc = frames_cls.detach().float().contiguous()
        content = torch.empty(b, d, dtype=f32, device=dev)
        gamma = torch.empty(b, d, dtype=f32, device=dev)
        beta = torch.empty(b, d, dtype=f32, device=dev)
        mod_op = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
        anchor = torch.empty(n, q, dtype=f32, device=dev)
        temp = torch.empty(n, d, dtype=f32, device=dev)
        be.template_fwd(v, fc, f2v, _operand(Wc.detach(), True), bc.detach(), _operand(Wg.detach(), True), bg.detach(),
                        _operand(Wb.detach(), True), bb.detach(), _operand(Wa.detach(), True), ba.detach(), content, gamma, beta,
                        mod_op, anchor, temp)
        ctx.save_for_backward(v, fc, gamma, beta, mod_op, anchor, Wc, Wg, Wb, Wa)
        ctx.extra = (bc, bg, bb, ba, f2v, vid_start)
        ctx.set_materialize_grads(False)
        return anchor, temp

    @staticmethod
    @once_differentiable
    def backward(ctx, g_anchor, g_temp):
        if g_anchor is None and g_temp is None:
            return (None,) * 12
        be = get_backend()
        v, fc, gamma, beta, mod_op, anchor, Wc, Wg, Wb, Wa = ctx.saved_tensors
        bc, bg, bb, ba, f2v, vid_start = ctx.extra
        n, d = fc.shape
        b, q = v.shape[0], Wa.shape[0]
        dev = fc.device
        f32 = torch.float32
        fix = lambda g: g if (g.is_contiguous() and g.dtype == f32) else g.contiguous().float()
        g_anchor = torch.zeros(n, q, dtype=f32, device=dev) if g_anchor is None else fix(g_anchor)
        g_temp = None if g_temp is None else fix(g_temp)
        params = (Wc, bc, Wg, bg, Wb, bb, Wa, ba)
        fused = _fuse_grads and all(p_.grad is not None and p_.grad.is_contiguous() for p_ in params)
        if fused:
            gs = [p_.grad for p_ in params]
        else:
            gs = [torch.zeros(p_.shape, dtype=f32, device=dev) for p_ in params]
        dpq_op = torch.empty(n, q, dtype=torch.bfloat16, device=dev)
        dmod = torch.empty(n, d, dtype=f32, device=dev)
        dpre = torch.empty(3, b, d, dtype=f32, device=dev)
        dfc = torch.empty(n, d, dtype=f32, device=dev)
        dv = torch.empty(b, d, dtype=f32, device=dev)
        be.template_bwd(g_anchor, g_temp, anchor, v, fc, f2v, vid_start, gamma, beta, mod_op, _operand(Wc.detach(), True),
                        _operand(Wg.detach(), True), _operand(Wb.detach(), True), _operand(Wa.detach(), True), dpq_op, dmod, dpre,
                        dfc, dv, gs[0], gs[1], gs[2], gs[3], gs[4], gs[5], gs[6], gs[7])
        out = (None,) * 8 if fused else tuple(gs)
        return (dv, dfc) + out + (None, None)


def template(videos_cls, frames_cls, Wc, bc, Wg, bg, Wb, bb, Wa, ba, f2v=None, vid_start=None):
    return TemplateFn.apply(videos_cls, frames_cls, Wc, bc, Wg, bg, Wb, bb, Wa, ba, f2v, vid_start)


class BoxHeadFn(Function):
    """(new_anchor, sine, sine_op) = the last Linear of ``bbox_embed`` on the bf16 hidden activation ``h``, the anchor refinement
    sigmoid(delta + inverse_sigmoid(anchor)) and the sine embedding of the (detached) refined anchor in ONE launch
    (query_decoder.py:205-219, 188-199; net_utils.py:29-63).  The sine outputs are constants for autograd, exactly as in the
    reference, where the next layer embeds ``new_reference_points.detach()``.  Backward: box_refine_bwd, then the data / weight
    gradients of the Linear as in LinearFn."""

    @staticmethod
    def forward(ctx, h, weight, bias, anchor, want_sine: bool, eps: float):
        be = get_backend()
        R, K = h.shape
        dev = h.device
        a = anchor.detach().float().contiguous()
        out = torch.empty(R, 4, dtype=torch.float32, device=dev)
        sine = torch.empty(R, 512, dtype=torch.float32, device=dev) if want_sine else None
        sine_op = torch.empty(R, 512, dtype=torch.bfloat16, device=dev) if want_sine else None
        hd = h.detach()
        be.box_head_fwd(hd, _operand(weight.detach(), True), bias.detach(), a, out, sine, sine_op, eps)
        ctx.save_for_backward(hd, weight, bias, out, a)
        ctx.eps = eps
        ctx.set_materialize_grads(False)
        if want_sine:
            ctx.mark_non_differentiable(sine, sine_op)
        return out, sine, sine_op

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _s, _so):
        if g is None:
            return (None,) * 6
        be = get_backend()
        hd, weight, bias, out, a = ctx.saved_tensors
        R, K = hd.shape
        g = g if (g.is_contiguous() and g.dtype == torch.float32) else g.contiguous().float()
        dd = torch.empty_like(out)
        da = torch.empty_like(out) if ctx.needs_input_grad[3] else None
        be.box_refine_bwd(out, a, g, dd, da, ctx.eps)
        dyo = _operand(dd)
        dh = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty(R, K, dtype=hd.dtype, device=hd.device)
            be.linear_bwd_data(dyo, _operand(weight.detach(), True), dh)
        dw, db = _wgrad(be, dyo, hd, weight, bias, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return dh, dw, db, da, None, None


def box_head(h, weight, bias, anchor, want_sine: bool = True, eps: float = 1e-3):
    return BoxHeadFn.apply(h, weight, bias, anchor, want_sine, eps)


class BoxMLPHeadFn(Function):
    """``bbox_embed`` (Linear-ReLU stack, net_utils.py:7-26) + anchor refinement + the next layer's sine embedding as ONE
    autograd node (query_decoder.py:205-219): forward = the hidden Linears (bf16 outputs) and ops.box_head's kernel; backward =
    stcat_box_head_bwd (refinement gradient, data gradient of the last Linear, ReLU mask) and one data-gradient GEMM per
    hidden layer with the ReLU mask of ITS input in the epilogue -- 3 dependent launches where the per-Linear nodes need 7
    (refine_bwd, cast, dgrad, relu_bwd, dgrad, relu_bwd, dgrad); weight / bias gradients go to the leaf streams."""

    @staticmethod
    def forward(ctx, x, x_op, anchor, want_sine, eps, nl, *wb):
        be = get_backend()
        ws, bs = wb[:nl], wb[nl:]
        xd = x.detach()
        xo = x_op.detach() if (x_op is not None and x_op.dtype == torch.bfloat16) else _cast_op(be, xd if xd.is_contiguous() else xd.contiguous())
        R = xo.shape[0]
        dev = xo.device
        acts = [xo]
        for i in range(nl - 1):
            hdn = torch.empty(R, ws[i].shape[0], dtype=torch.bfloat16, device=dev)
            be.linear_fwd(acts[-1], _operand(ws[i].detach(), True), bs[i].detach(), hdn, relu=True)
            acts.append(hdn)
        a = anchor.detach().float().contiguous()
        out = torch.empty(R, 4, dtype=torch.float32, device=dev)
        sine = torch.empty(R, 512, dtype=torch.float32, device=dev) if want_sine else None
        sine_op = torch.empty(R, 512, dtype=torch.bfloat16, device=dev) if want_sine else None
        be.box_head_fwd(acts[-1], _operand(ws[-1].detach(), True), bs[-1].detach(), a, out, sine, sine_op, eps)
        ctx.save_for_backward(out, a, *acts, *ws)
        ctx.biases = bs
        ctx.nl, ctx.eps = nl, eps
        ctx.x_dtype = x.dtype
        ctx.set_materialize_grads(False)
        if want_sine:
            ctx.mark_non_differentiable(sine, sine_op)
        return out, sine, sine_op

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _s, _so):
        nl = ctx.nl
        if g is None:
            return (None,) * (6 + 2 * nl)
        be = get_backend()
        saved = ctx.saved_tensors
        out, a = saved[0], saved[1]
        acts, ws = saved[2:2 + nl], saved[2 + nl:]
        bs = ctx.biases
        R = out.shape[0]
        dev = out.device
        g = g if (g.is_contiguous() and g.dtype == torch.float32) else g.contiguous().float()
        need = ctx.needs_input_grad
        da = torch.empty_like(out) if need[2] else None
        dd_op = torch.empty(R, 4, dtype=torch.bfloat16, device=dev)
        cur = torch.empty(R, ws[-1].shape[1], dtype=torch.bfloat16, device=dev)
        be.box_head_bwd(g, out, a, _operand(ws[-1].detach(), True), acts[-1], dd_op, cur, da, ctx.eps)
        dws, dbs = [None] * nl, [None] * nl
        dws[-1], dbs[-1] = _wgrad(be, dd_op, acts[-1], ws[-1], bs[-1], need[6 + nl - 1], need[6 + 2 * nl - 1])
        dx = None
        for i in range(nl - 2, -1, -1):  # cur = gradient w.r.t. the pre-activation of hidden layer i (bf16 [R, N_i])
            dws[i], dbs[i] = _wgrad(be, cur, acts[i], ws[i], bs[i], need[6 + i], need[6 + nl + i])
            wo = _operand(ws[i].detach(), True)
            if i > 0:
                prev = torch.empty(R, ws[i].shape[1], dtype=torch.bfloat16, device=dev)
                be.linear_bwd_data(cur, wo, prev, relu_y=acts[i])  # acts[i] = ReLU output of hidden layer i - 1: mask in the epilogue
                cur = prev
            elif need[0]:
                dx = torch.empty(R, ws[0].shape[1], dtype=ctx.x_dtype, device=dev)
                be.linear_bwd_data(cur, wo, dx)
        return (dx, None, da, None, None, None, *dws, *dbs)


def box_mlp_head(mlp_layers, x, x_op, anchor, want_sine: bool = True, eps: float = 1e-3):
    """(refined anchor [R, 4], sine [R, 512] or None, its bf16 copy or None) from the query activations ``x`` [R, d]"""
    ws = [l.weight for l in mlp_layers]
    bs = [l.bias for l in mlp_layers]
    return BoxMLPHeadFn.apply(x, x_op, anchor, want_sine, eps, len(ws), *ws, *bs)


class MulOperandFn(Function):
    """(a[:, :c] * b, its bf16 GEMM-operand copy) in one launch (query_sine = sine[..., :d] * query_scale(out),
    query_decoder.py:196-199); ``a`` is a constant (the sine embedding of a detached anchor).  Backward: d b = g * a[:, :c]."""

    @staticmethod
    def forward(ctx, a, b, c):
        be = get_backend()
        bd = b.detach().float().contiguous()
        ad = a.detach()
        out = torch.empty(bd.shape, dtype=torch.float32, device=bd.device)
        out_op = torch.empty(bd.shape, dtype=torch.bfloat16, device=bd.device)
        c_op = None
        if c is not None and c.dtype == torch.float32 and c.shape == bd.shape and c.is_contiguous():
            c_op = torch.empty(bd.shape, dtype=torch.bfloat16, device=bd.device)
            be.mul_cast(ad, bd, out, out_op, c.detach(), c_op)
        else:
            be.mul_cast(ad, bd, out, out_op)
        ctx.save_for_backward(ad)
        ctx.shape = bd.shape
        ctx.mark_non_differentiable(*([out_op] if c_op is None else [out_op, c_op]))
        ctx.set_materialize_grads(False)
        return out, out_op, c_op

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _unused, _unused2):
        if g is None:
            return None, None, None
        (ad,) = ctx.saved_tensors
        g = g if g.is_contiguous() else g.contiguous()
        db = torch.empty(ctx.shape, dtype=torch.float32, device=g.device)
        get_backend().mul_cast_bwd(g, ad, db)
        return None, db, None


def mul_operand(a, b, c=None):
    """a fp32 [R, >= k] (row-major, constant), b fp32 [R, k] -> (a[:, :k] * b fp32, its bf16 copy, bf16(c) or None): ``c``
    (fp32 [R, k], optional) is a second tensor whose GEMM-operand copy rides in the same launch"""
    assert not a.requires_grad
    return MulOperandFn.apply(a, b, c)


def sted_score(pred_sted: torch.Tensor, durations, return_map: bool = False):
    """Temporal start/end scoring (post_processor.py:30-53).  Returns (best flat index [b] int32 on
    device, score map or None)."""
    be = get_backend()
    b, t, _ = pred_sted.shape
    dur = torch.as_tensor(list(durations), dtype=torch.int32).to(pred_sted.device, non_blocking=True)
    score = torch.empty(b, t, t, dtype=torch.float32, device=pred_sted.device) if return_map else None
    best = torch.empty(b, dtype=torch.int32, device=pred_sted.device)
    be.sted_score(pred_sted.detach().float().contiguous(), dur, score, best)
    return best, score

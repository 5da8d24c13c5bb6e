"""TEST INFRASTRUCTURE ONLY: a torch-CPU emulation of ``stcat_b200.cabi.CudaBackend``.

It implements the *contract* of every C-ABI entry point (include/stcat_b200.h) with plain torch ops so
that the host-side composition (stcat_b200/{ops,encoder,decoder,pipeline}.py: layouts, index maps,
hand-written backward chains) can be checked against the oracle in the GPU-less build container.
It is installed only by tests through ``ops.set_backend``; product code never selects it and
``CudaBackend`` raises on CPU tensors.
"""
import torch


def _f(t):
    return t.float() if t.dtype != torch.float32 else t


def _store(dst, val):
    dst.copy_(val.to(dst.dtype))


_M64 = (1 << 64) - 1


def _s64(v):
    """python int (mod 2^64) -> the int64 with the same bit pattern"""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z, k):
    """logical shift right of an int64 tensor (torch's >> is arithmetic)"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def drop_keep_scale(n, p, seed, offset, device="cpu"):
    """Bit-exact restatement of drop_bits24 / make_drop (csrc/common.cuh): (keep mask [n] bool, scale)."""
    t = int(float(torch.tensor(p, dtype=torch.float32)) * 16777216.0)  # the C ABI takes p as a float
    t = min(t, 16777215)
    idx = torch.arange(n, dtype=torch.int64, device=device)
    z = idx + _s64(offset + seed * 0x9E3779B97F4A7C15)  # 64-bit counter (wraps like the device's uint64 arithmetic)
    m32 = 0xFFFFFFFF
    lo, hi = z & m32, _lsr(z, 32)
    x = lo ^ ((hi * 0x9E3779B1) & m32)  # uint32 arithmetic carried in int64 lanes: every product stays below 2^64
    x = x ^ (x >> 16)
    x = (x * 0x21F0AAAD) & m32
    x = x ^ (x >> 15)
    x = (x * 0x735A2D97) & m32
    x = x ^ (x >> 15)
    bits = x >> 8
    return bits >= t, float(16777216.0 / (16777216.0 - t))


class LayoutError(AssertionError):
    pass


def _mat(t, name):
    """the layout rule of CudaBackend._mat (2-D view with unit column stride), minus the device check: a view the C ABI
    would reject must fail the CPU tests of the host composition too"""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise LayoutError(f"{name}: expected a 2-D row-major view, got shape {tuple(t.shape)} stride {t.stride()}")
    if t.dtype not in (torch.float32, torch.bfloat16):
        raise LayoutError(f"{name}: unsupported dtype {t.dtype}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _flat(t, name, dtype=None):
    """CudaBackend._flat: contiguous, of the expected dtype (None passes)"""
    if t is None:
        return
    if not t.is_contiguous():
        raise LayoutError(f"{name}: must be contiguous, got shape {tuple(t.shape)} stride {t.stride()}")
    if dtype is not None and t.dtype != dtype:
        raise LayoutError(f"{name}: expected {dtype}, got {t.dtype}")


F32, BF16 = torch.float32, torch.bfloat16


class EmuBackend:
    name = "emu"

    def __init__(self):
        self.launches = 0

    # -- linear --------------------------------------------------------
    def linear_fwd(self, x, w, bias, y, relu=False, accumulate=False):
        _mat(x, "x"), _mat(w, "w"), _mat(y, "y"), _flat(bias, "bias", F32)
        assert w.shape[1] == x.shape[1] and tuple(y.shape) == (x.shape[0], w.shape[0])
        v = _f(x) @ _f(w).t()
        if bias is not None:
            v = v + bias
        if accumulate:
            v = v + _f(y)
        if relu:
            v = v.relu()
        _store(y, v)
        self.launches += 1

    def linear_dropout_fwd(self, x, w, bias, y, relu, drop):
        _mat(x, "x"), _mat(w, "w"), _mat(y, "y"), _flat(bias, "bias", F32)
        assert w.shape[1] == x.shape[1] and tuple(y.shape) == (x.shape[0], w.shape[0]) and y.is_contiguous()
        v = _f(x) @ _f(w).t()
        if bias is not None:
            v = v + bias
        if relu:
            v = v.relu()
        keep, sc = drop_keep_scale(v.numel(), drop[0], drop[1], drop[2], v.device)
        _store(y, v * (keep.reshape(v.shape).float() * sc))
        self.launches += 1

    def linear_bwd_data(self, dy, w, dx, accumulate=False, relu_y=None, dbias=None, alpha=1.0):
        _mat(dy, "dy"), _mat(w, "w"), _mat(dx, "dx"), _flat(dbias, "dbias", F32)
        assert alpha == 1.0 or (relu_y is not None and not accumulate)
        assert w.shape[0] == dy.shape[1] and tuple(dx.shape) == (dy.shape[0], w.shape[1])
        if relu_y is not None:
            _mat(relu_y, "relu_y")
            assert relu_y.shape == dx.shape
        v = _f(dy) @ _f(w)
        if accumulate:
            v = v + _f(dx)
        if alpha != 1.0:
            v = v * alpha
        if relu_y is not None:
            v = v * (_f(relu_y) > 0)
        _store(dx, v)
        if dbias is not None:
            dbias.add_(_f(dx).sum(0))
        self.launches += 1

    def linear_bwd_weight(self, dy, x, dw, db, accumulate=False):
        _mat(dy, "dy"), _mat(x, "x"), _mat(dw, "dw"), _flat(db, "db", F32)
        assert x.shape[0] == dy.shape[0] and tuple(dw.shape) == (dy.shape[1], x.shape[1]) and dw.dtype == F32
        v = _f(dy).t() @ _f(x)
        s = _f(dy).sum(0)
        if accumulate:
            v = v + dw
            if db is not None:
                s = s + db
        _store(dw, v)
        if db is not None:
            _store(db, s)
        self.launches += 1

    def linear_group(self, kind, jobs):
        assert 1 <= len(jobs) <= 12
        in_dt = None
        for job in jobs:
            out = job["out"]
            _mat(out, "out")
            rows, cols = out.shape
            assert 1 <= len(job["terms"]) <= 3
            _flat(job.get("dbias"), "dbias", F32)
            for a, b, bias in job["terms"]:
                _mat(a, "a"), _mat(b, "b"), _flat(bias, "bias", F32)
                assert a.dtype == b.dtype and (in_dt is None or in_dt == a.dtype), "operands of a grouped launch share one dtype"
                in_dt = a.dtype
                if kind == 0:
                    assert a.shape[0] == rows and tuple(b.shape) == (cols, a.shape[1])
                elif kind == 1:
                    assert a.shape[0] == rows and tuple(b.shape) == (a.shape[1], cols)
                else:
                    assert tuple(a.shape) == (a.shape[0], rows) and tuple(b.shape) == (a.shape[0], cols)
            v = _f(out) if job.get("accumulate") else None
            for a, b, bias in job["terms"]:
                if kind == 0:
                    t = _f(a) @ _f(b).t()
                    if bias is not None:
                        t = t + bias
                elif kind == 1:
                    t = _f(a) @ _f(b)
                else:
                    t = _f(a).t() @ _f(b)
                v = t if v is None else v + t
            if job.get("relu"):
                v = v.relu()
            _store(out, v)
            if kind == 2 and job.get("dbias") is not None:
                job["dbias"].add_(_f(job["terms"][0][0]).sum(0))
        self.launches += 1

    # -- layernorm -----------------------------------------------------
    @staticmethod
    def _ln_mask(x, drop):
        """keep-mask * scale of the fused dropout (element index row * d + column), or None"""
        if drop is None or drop[0] <= 0:
            return None
        keep, sc = drop_keep_scale(x.numel(), drop[0], drop[1], drop[2], x.device)
        return keep.reshape(x.shape).float() * sc

    def layernorm_fwd(self, x, res, gamma, beta, y, y_bf16, mean, rstd, eps=1e-5, drop=None):
        for t, nm in ((x, "x"), (res, "res"), (gamma, "gamma"), (beta, "beta"), (y, "y"), (mean, "mean"), (rstd, "rstd")):
            _flat(t, nm, F32)
        _flat(y_bf16, "y_bf16", BF16)
        mk = self._ln_mask(x, drop)
        x = x if mk is None else x * mk
        z = x if res is None else x + res
        mu = z.mean(-1)
        var = ((z - mu[:, None]) ** 2).mean(-1)
        rs = torch.rsqrt(var + eps)
        out = (z - mu[:, None]) * rs[:, None] * gamma + beta
        y.copy_(out)
        if y_bf16 is not None:
            y_bf16.copy_(out.to(torch.bfloat16))
        mean.copy_(mu)
        rstd.copy_(rs)
        self.launches += 1

    def layernorm_bwd(self, dy, x, res, gamma, mean, rstd, dz, dgamma, dbeta, dz_bf16=None, dbias=None, drop=None):
        for t, nm in ((dy, "dy"), (x, "x"), (res, "res"), (gamma, "gamma"), (mean, "mean"), (rstd, "rstd"), (dz, "dz"),
                      (dgamma, "dgamma"), (dbeta, "dbeta"), (dbias, "dbias")):
            _flat(t, nm, F32)
        _flat(dz_bf16, "dz_bf16", BF16)
        mk = self._ln_mask(x, drop)
        x = x if mk is None else x * mk
        z = x if res is None else x + res
        xh = (z - mean[:, None]) * rstd[:, None]
        dg = dy * gamma
        s1 = dg.mean(-1, keepdim=True)
        s2 = (dg * xh).mean(-1, keepdim=True)
        dz.copy_(rstd[:, None] * (dg - s1 - xh * s2))
        dgamma.add_((dy * xh).sum(0))
        dbeta.add_(dy.sum(0))
        dzx = dz if mk is None else dz * mk  # gradient w.r.t. x (dz itself stays the residual branch's gradient)
        if dz_bf16 is not None:
            dz_bf16.copy_(dzx.to(torch.bfloat16))
        if dbias is not None:
            dbias.add_(dzx.sum(0))
        self.launches += 1

    # -- attention -----------------------------------------------------
    @staticmethod
    def _scores(q1, q2, k1, k2, key_mask, B, H, Lq, Lk, scale):
        hd = lambda t, L: _f(t).reshape(B, L, H, 32).permute(0, 2, 1, 3)
        s = hd(q1, Lq) @ hd(k1, Lk).transpose(-1, -2)
        if q2 is not None:
            s = s + hd(q2, Lq) @ hd(k2, Lk).transpose(-1, -2)
        s = s * scale
        if key_mask is not None:
            s = s.masked_fill(key_mask.bool()[:, None, None, :], float("-inf"))
        return s, hd

    def dropout(self, x, out, p, seed, offset):
        assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype and x.shape == out.shape
        keep, sc = drop_keep_scale(x.numel(), p, seed, offset, x.device)
        _store(out, (_f(x).reshape(-1) * keep * sc).reshape(x.shape))
        self.launches += 1

    @staticmethod
    def _pmask(drop, B, H, Lq, Lk, device):
        if drop is None or drop[0] <= 0:
            return None
        keep, sc = drop_keep_scale(B * H * Lq * Lk, drop[0], drop[1], drop[2], device)
        return keep.reshape(B, H, Lq, Lk).float() * sc

    @staticmethod
    def _attn_layout(q1, q2, k1, k2, B, H, Lq, Lk, others):
        ldq, ldk = _mat(q1, "q1"), _mat(k1, "k1")
        assert q1.shape == (B * Lq, H * 32) and k1.shape == (B * Lk, H * 32)
        assert (q2 is None) == (k2 is None)
        if q2 is not None:  # the second score part shares the leading dimensions of the first
            assert _mat(q2, "q2") == ldq and _mat(k2, "k2") == ldk and q2.shape == q1.shape and k2.shape == k1.shape
        for t, nm, L in others:
            _mat(t, nm)
            assert t.shape == (B * L, H * 32) and t.dtype == q1.dtype, nm

    @staticmethod
    def _variant(q1, q2, want_pavg, B, H, Lq, Lk):
        """the kernel the C ABI would run for bf16 operands, as far as rounding goes (oracle.kernel_attention_variant)"""
        if q1.dtype != BF16:
            return "f32p"
        from oracle.stcat_oracle import kernel_attention_variant

        return kernel_attention_variant(B, H, Lq, Lk, q2 is not None, want_pavg)

    def dropout_bits(self, bits, cols, p, seed, offset):
        """the keep mask of a [rows, cols] site packed one bit per element (stcat_dropout_bits)"""
        rows, wpr = bits.shape
        assert bits.dtype == torch.int32 and wpr * 32 >= cols
        keep, _ = drop_keep_scale(rows * cols, p, seed, offset, bits.device)
        k = torch.zeros(rows, wpr * 32, dtype=torch.int64, device=bits.device)
        k[:, :cols] = keep.reshape(rows, cols).to(torch.int64)
        w = (k.view(rows, wpr, 32) << torch.arange(32, dtype=torch.int64, device=bits.device)).sum(-1)
        bits.copy_(torch.where(w >= (1 << 31), w - (1 << 32), w).to(torch.int32))
        self.launches += 1

    def attention_fwd(self, q1, q2, k1, k2, v, o, key_mask, lse, p_avg, B, H, Lq, Lk, scale, drop=None, bits=None):
        self._attn_layout(q1, q2, k1, k2, B, H, Lq, Lk, ((v, "v", Lk), (o, "o", Lq)))
        _flat(key_mask, "key_mask", torch.uint8), _flat(lse, "lse", F32), _flat(p_avg, "p_avg", F32)
        s, hd = self._scores(q1, q2, k1, k2, key_mask, B, H, Lq, Lk, scale)
        p = torch.softmax(s, -1)
        m = self._pmask(drop, B, H, Lq, Lk, p.device)
        variant = self._variant(q1, q2, p_avg is not None, B, H, Lq, Lk)
        rb = lambda t: t.to(BF16).float()
        if variant == "tc":
            # un-normalised probabilities relative to the running row max, rounded to bf16 per 256-key tile; fp32 normaliser
            from oracle.stcat_oracle import TC_KEY_TILE

            vh = hd(v, Lk)
            acc = den = m_run = None
            for k0 in range(0, Lk, TC_KEY_TILE):
                st = s[..., k0:k0 + TC_KEY_TILE]
                m_new = st.max(-1, keepdim=True)[0] if m_run is None else torch.maximum(m_run, st.max(-1, keepdim=True)[0])
                m_safe = torch.where(torch.isinf(m_new), torch.zeros_like(m_new), m_new)
                e = torch.exp(st - m_safe)
                em = e if m is None else e * m[..., k0:k0 + TC_KEY_TILE]
                part = rb(em) @ vh[:, :, k0:k0 + TC_KEY_TILE]
                if acc is None:
                    acc, den = part, e.sum(-1, keepdim=True)
                else:
                    alpha = torch.where(torch.isinf(m_run), torch.zeros_like(m_run), torch.exp(m_run - m_safe))
                    acc, den = acc * alpha + part, den * alpha + e.sum(-1, keepdim=True)
                m_run = m_new
            out = acc / den.clamp_min(1e-30)
            if m is not None:
                p = p * m
        else:
            if m is not None:
                p = p * m
            out = (rb(p) if variant == "mma" else p) @ hd(v, Lk)
        out = out.permute(0, 2, 1, 3).reshape(B * Lq, H * 32)
        _store(o, out)
        lse.copy_(torch.logsumexp(s, -1))
        if p_avg is not None:
            p_avg.add_(p.mean(1))
        self.launches += 1

    def attention_bwd(self, q1, q2, k1, k2, v, d_o, key_mask, lse, dp_avg, delta, dq1, dq2, dk1, dk2, dv, B, H, Lq, Lk,
                      scale, o=None, drop=None, bits=None):
        oth = [(v, "v", Lk), (d_o, "d_o", Lq), (dq1, "dq1", Lq), (dk1, "dk1", Lk), (dv, "dv", Lk)] + ([(o, "o", Lq)] if o is not None else [])
        self._attn_layout(q1, q2, k1, k2, B, H, Lq, Lk, oth)
        if q2 is not None:
            assert _mat(dq2, "dq2") == _mat(dq1, "dq1") and _mat(dk2, "dk2") == _mat(dk1, "dk1")
        _flat(key_mask, "key_mask", torch.uint8), _flat(lse, "lse", F32), _flat(dp_avg, "dp_avg", F32), _flat(delta, "delta", F32)
        s, hd = self._scores(q1, q2, k1, k2, key_mask, B, H, Lq, Lk, scale)
        p = torch.exp(s - lse[..., None])
        g = hd(d_o, Lq)
        dp = g @ hd(v, Lk).transpose(-1, -2)
        if dp_avg is not None:
            dp = dp + dp_avg[:, None] / H
        m = self._pmask(drop, B, H, Lq, Lk, p.device)
        pm = p
        if m is not None:  # o and p_avg were formed from p * m
            dp = dp * m
            pm = p * m
        variant = self._variant(q1, q2, dp_avg is not None, B, H, Lq, Lk)
        if variant == "tc" and o is not None:
            dl = (hd(o, Lq) * g).sum(-1, keepdim=True)  # the tcgen05 backward takes delta = dO . O from the stored bf16 output
        else:
            dl = (p * dp).sum(-1, keepdim=True)
        delta.copy_(dl.squeeze(-1))
        ds = p * (dp - dl) * scale
        if variant in ("tc", "mma"):  # P and dS are rounded to bf16 for the gradient products (attention_tc.cu, attention_small_mma.cu)
            pm, ds = pm.to(BF16).float(), ds.to(BF16).float()
        un = lambda t, L: t.permute(0, 2, 1, 3).reshape(B * L, H * 32)
        _store(dv, un(pm.transpose(-1, -2) @ g, Lk))
        _store(dq1, un(ds @ hd(k1, Lk), Lq))
        _store(dk1, un(ds.transpose(-1, -2) @ hd(q1, Lq), Lk))
        if q2 is not None:
            _store(dq2, un(ds @ hd(k2, Lk), Lq))
            _store(dk2, un(ds.transpose(-1, -2) @ hd(q2, Lq), Lk))
        self.launches += 2

    # -- element-wise --------------------------------------------------
    def add(self, a, b, out, out_bf16=None):
        _flat(a, "a", F32), _flat(b, "b", F32), _flat(out, "out", F32), _flat(out_bf16, "out_bf16", BF16)
        z = a + b
        if out is not None:
            out.copy_(z)
        if out_bf16 is not None:
            out_bf16.copy_(z.to(torch.bfloat16))
        self.launches += 1

    def relu_bwd(self, y, dy):
        _flat(y, "y"), _flat(dy, "dy")
        dy.masked_fill_(~(_f(y) > 0), 0)
        self.launches += 1

    def cast_bf16(self, x, out, transpose=False):
        _flat(x, "x", F32), _flat(out, "out", BF16)
        out.copy_((x.t() if transpose else x).to(torch.bfloat16))
        self.launches += 1

    # -- optimizer-side step ---------------------------------------------
    def sumsq(self, x, accum):
        _flat(x, "x", F32), _flat(accum, "accum", F32)
        accum.add_((x.double() ** 2).sum().float())
        self.launches += 1

    def adamw_step(self, p, g, m, v, ema, shadow, lr, beta1, beta2, eps, weight_decay, step, total_sumsq, max_norm, ema_decay):
        """the per-element order of stcat_adamw_step (include/stcat_b200.h), in fp32"""
        for t, nm in ((p, "p"), (g, "g"), (m, "m"), (v, "v"), (ema, "ema"), (total_sumsq, "total_sumsq")):
            _flat(t, nm, F32)
        _flat(shadow, "shadow", BF16)
        g = g.clone()
        if max_norm > 0:
            g.mul_(torch.clamp(max_norm / (total_sumsq.sqrt() + 1e-6), max=1.0))
        bc1 = 1.0 - beta1 ** step
        bc2s = (1.0 - beta2 ** step) ** 0.5
        p.mul_(1 - lr * weight_decay)
        m.add_((g - m) * (1 - beta1))
        v.mul_(beta2).add_((1 - beta2) * g * g)
        p.sub_((lr / bc1) * (m / (v.sqrt() / bc2s + eps)))
        if ema is not None:
            ema.mul_(ema_decay).add_((1 - ema_decay) * p)
        if shadow is not None:
            shadow.copy_(p.to(torch.bfloat16))
        self.launches += 1

    def sted_score(self, sted, durations, score, best):
        b, t, _ = sted.shape
        ls = torch.log_softmax(sted[:, :, 0], 1)
        le = torch.log_softmax(sted[:, :, 1], 1)
        ii = torch.arange(t)[:, None]
        jj = torch.arange(t)[None, :]
        for v in range(b):
            dur = int(durations[v])
            pen = torch.zeros(t, t)
            pen[(jj <= ii) | (ii >= dur) | (jj >= dur)] = -1e32
            sc = pen + (ls[v][:, None] + le[v][None, :])
            if score is not None:
                score[v].copy_(sc)
            best[v] = int(sc.flatten().argmax())
        self.launches += 1

    def set_gemm_sm_limit(self, n):
        pass

    def set_sm_cap(self, n):
        pass

    def set_dropout_step(self, counter):
        pass

    def pos_sine(self, mask, out, num_pos_feats, temperature, scale):
        import math

        nm = ~mask.bool()
        y = nm.cumsum(1, dtype=torch.float32)
        x = nm.cumsum(2, dtype=torch.float32)
        y = y / (y[:, -1:, :] + 1e-6) * scale
        x = x / (x[:, :, -1:] + 1e-6) * scale
        k = torch.arange(num_pos_feats, dtype=torch.float32)
        dim_t = temperature ** (2 * torch.div(k, 2, rounding_mode="floor") / num_pos_feats)
        px, py = x[..., None] / dim_t, y[..., None] / dim_t
        px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), 4).flatten(3)
        py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), 4).flatten(3)
        out.copy_(torch.cat((py, px), 3))
        self.launches += 1

    def box_interp(self, frame_ids, boxes, out, first):
        ids = [int(v) for v in frame_ids]
        for i in range(out.shape[0]):
            fid = first + i
            if not ids or fid < ids[0] or fid > ids[-1]:
                out[i] = -1.0
                continue
            lo = max(j for j, v in enumerate(ids) if v <= fid)
            nx = min(lo + 1, len(ids) - 1)
            span = ids[nx] - ids[lo]
            t = (fid - ids[lo]) / span if span > 0 else 0.0
            out[i] = boxes[lo] + t * (boxes[nx] - boxes[lo])
        self.launches += 1

    def map2d_pool(self, x, valid, out):
        B, N, d = x.shape
        out.zero_()
        for i in range(N):
            run = torch.full((B, d), float("-inf"))
            for j in range(i, N):
                run = torch.maximum(run, x[:, j])
                if valid.view(N, N)[i, j]:
                    out[:, :, i, j] = run
        self.launches += 1

    # -- layout glue (csrc/assembly.cu) ----------------------------------
    def token_assembly(self, vis, vpos, text, f2v, frame_cls, local_pos, X, POS, qk_op, x_op):
        for t, nm in ((vis, "vis"), (vpos, "vpos"), (text, "text"), (frame_cls, "frame_cls"), (local_pos, "local_pos"), (X, "X"),
                      (POS, "POS")):
            _flat(t, nm, F32)
        _flat(qk_op, "qk_op", BF16), _flat(x_op, "x_op", BF16), _flat(f2v, "f2v", torch.int64)
        n, d = vis.shape[0], vis.shape[1]
        L, b = text.shape[0], text.shape[1]
        assert b == 1 or f2v is not None
        xv = vis.flatten(2).transpose(1, 2)
        pv = vpos.flatten(2).transpose(1, 2)
        xt = text.transpose(0, 1)
        xt = xt.expand(n, L, d) if b == 1 else xt.index_select(0, f2v)
        X.copy_(torch.cat([frame_cls.view(1, 1, d).expand(n, 1, d), xv, xt], 1))
        POS.copy_(torch.cat([local_pos.view(1, 1, d).expand(n, 1, d), pv, pv.new_zeros(n, L, d)], 1))
        if qk_op is not None:
            _store(qk_op, X + POS)
        if x_op is not None:
            _store(x_op, X)
        self.launches += 1

    def token_assembly_bwd(self, dX, dvis, dtext, dcls, vid_start, HW, L, b):
        _flat(dX, "dX", F32), _flat(dvis, "dvis", F32), _flat(dtext, "dtext", F32), _flat(dcls, "dcls", F32)
        _flat(vid_start, "vid_start", torch.int64)
        n, S, d = dX.shape
        assert S == 1 + HW + L and (b == 1 or vid_start is not None)
        if dvis is not None:
            dvis.copy_(dX[:, 1:1 + HW].transpose(1, 2).reshape(dvis.shape))
        if dtext is not None and L > 0:
            st = [0, n] if b == 1 else [int(v) for v in vid_start]
            for v in range(b):
                dtext[:, v] = dX[st[v]:st[v + 1], 1 + HW:].sum(0)
        if dcls is not None:
            dcls.copy_(dX[:, 0].sum(0).view(dcls.shape))
        self.launches += 1

    def mem_operands(self, X, POS, mem_op, pos_op, mempos_op, cls):
        _flat(X, "X", F32), _flat(POS, "POS", F32), _flat(mem_op, "mem_op", BF16), _flat(pos_op, "pos_op", BF16)
        _flat(mempos_op, "mempos_op", BF16), _flat(cls, "cls", F32)
        n, S, d = X.shape
        _store(mem_op, X[:, 1:].reshape(n * (S - 1), d))
        if pos_op is not None:
            _store(pos_op, POS[:, 1:].reshape(n * (S - 1), d))
        if mempos_op is not None:
            _store(mempos_op, (X[:, 1:] + POS[:, 1:]).reshape(n * (S - 1), d))
        if cls is not None:
            cls.copy_(X[:, 0])
        self.launches += 1

    def mem_operands_bwd(self, g_mem, g_mempos, g_cls, dX):
        _flat(g_mem, "g_mem"), _flat(g_mempos, "g_mempos"), _flat(g_cls, "g_cls", F32), _flat(dX, "dX", F32)
        n, S, d = dX.shape
        dX.zero_()
        if g_cls is not None:
            dX[:, 0] = g_cls
        if g_mem is not None:
            dX[:, 1:] += _f(g_mem).view(n, S - 1, d)
        if g_mempos is not None:
            dX[:, 1:] += _f(g_mempos).view(n, S - 1, d)
        self.launches += 1

    def template_fwd(self, videos_cls, frames_cls, f2v, Wc, bc, Wg, bg, Wb, bb, Wa, ba, content, gamma, beta, mod_op, anchor, temp_query):
        for t, nm in ((Wc, "Wc"), (Wg, "Wg"), (Wb, "Wb"), (Wa, "Wa"), (mod_op, "mod_op")):
            _flat(t, nm, BF16)
        for t, nm in ((videos_cls, "videos_cls"), (frames_cls, "frames_cls"), (bc, "bc"), (bg, "bg"), (bb, "bb"), (ba, "ba"),
                      (content, "content"), (gamma, "gamma"), (beta, "beta"), (anchor, "anchor")):
            _flat(t, nm, F32)
        b = videos_cls.shape[0]
        assert b == 1 or f2v is not None
        v = videos_cls.to(BF16).float()
        content.copy_(v @ _f(Wc).t() + bc)
        gamma.copy_(torch.tanh(v @ _f(Wg).t() + bg))
        beta.copy_(torch.tanh(v @ _f(Wb).t() + bb))
        g, bt = (gamma, beta) if b == 1 else (gamma.index_select(0, f2v), beta.index_select(0, f2v))
        _store(mod_op, g * frames_cls + bt)
        anchor.copy_(torch.sigmoid(_f(mod_op) @ _f(Wa).t() + ba))
        if temp_query is not None:
            _flat(temp_query, "temp_query", F32)
            temp_query.copy_(content.expand(frames_cls.shape[0], -1) if b == 1 else content.index_select(0, f2v))
        self.launches += 2

    def template_bwd(self, g_anchor, g_temp, anchor, videos_cls, frames_cls, f2v, vid_start, gamma, beta, mod_op, Wc, Wg, Wb, Wa,
                     dpq_op, dmod, dpre, d_frames_cls, d_videos_cls, dWc, dbc, dWg, dbg, dWb, dbb, dWa, dba):
        n, d = frames_cls.shape
        b = videos_cls.shape[0]
        assert b == 1 or (f2v is not None and vid_start is not None)
        _store(dpq_op, g_anchor * anchor * (1 - anchor))
        dmod.copy_(_f(dpq_op) @ _f(Wa))
        g = gamma if b == 1 else gamma.index_select(0, f2v)
        d_frames_cls.copy_(dmod * g)
        st = [0, n] if b == 1 else [int(v) for v in vid_start]
        for v in range(b):
            sl = slice(st[v], st[v + 1])
            dpre[0, v] = g_temp[sl].sum(0) if g_temp is not None else 0.0
            dpre[1, v] = (dmod[sl] * frames_cls[sl]).sum(0) * (1 - gamma[v] ** 2)
            dpre[2, v] = dmod[sl].sum(0) * (1 - beta[v] ** 2)
        dWa += _f(dpq_op).t() @ _f(mod_op)
        dba += _f(dpq_op).sum(0)
        vb = videos_cls.to(BF16).float()
        dv = torch.zeros_like(d_videos_cls)
        for which, (W, dW, db) in enumerate(((Wc, dWc, dbc), (Wg, dWg, dbg), (Wb, dWb, dbb))):
            dp = dpre[which].to(BF16).float()
            dv += dp @ _f(W)
            if dW is not None:
                dW += dp.t() @ vb
            if db is not None:
                db += dp.sum(0)
        d_videos_cls.copy_(dv)
        self.launches += 3

    def box_head_fwd(self, h, W, bias, anchor, out, sine, sine_op, eps=1e-3):
        import math

        _mat(h, "h"), _flat(W, "W", BF16), _flat(bias, "bias", F32), _flat(anchor, "anchor", F32), _flat(out, "out", F32)
        _flat(sine, "sine", F32), _flat(sine_op, "sine_op", BF16)
        assert h.dtype == BF16 and W.shape[0] == 4
        delta = _f(h) @ _f(W).t() + bias
        x = anchor.clamp(0, 1)
        out.copy_(torch.sigmoid(delta + torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))))
        if sine is not None:
            k = torch.arange(128, dtype=torch.float32)
            p = (out * (2 * math.pi))[..., None] / (10000 ** (2 * torch.div(k, 2, rounding_mode="floor") / 128))
            e = torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)
            sine.copy_(torch.cat((e[..., 1, :], e[..., 0, :], e[..., 2, :], e[..., 3, :]), dim=-1))
            if sine_op is not None:
                _store(sine_op, sine)
        self.launches += 1

    def mul_cast(self, a, b, out_f32, out_bf16, c_in=None, c_out=None):
        _mat(a, "a"), _flat(b, "b", F32), _flat(out_f32, "out_f32", F32), _flat(out_bf16, "out", BF16)
        v = a[:, : b.shape[1]] * b
        if out_f32 is not None:
            out_f32.copy_(v)
        _store(out_bf16, v)
        if c_out is not None:
            _flat(c_in, "c_in", F32), _flat(c_out, "c_out", BF16)
            assert c_in.shape == b.shape
            _store(c_out, c_in)
        self.launches += 1

    def mul_cast_bwd(self, g, a, db):
        _mat(a, "a"), _flat(g, "g"), _flat(db, "db", F32)
        db.copy_(_f(g) * a[:, : db.shape[1]])
        self.launches += 1

    def box_refine_bwd(self, out, anchor, g, ddelta, danchor, eps=1e-3):
        dz = g * out * (1 - out)
        ddelta.copy_(dz)
        if danchor is not None:
            x = anchor
            d = torch.zeros_like(x)
            inside = (x > 0) & (x < 1)
            d = d + torch.where(inside & (x > eps), 1.0 / x.clamp(min=1e-30), torch.zeros_like(x))
            d = d + torch.where(inside & (1 - x > eps), 1.0 / (1 - x).clamp(min=1e-30), torch.zeros_like(x))
            danchor.copy_(dz * d)
        self.launches += 1

    def box_head_bwd(self, g, out, anchor, W, h, dd_op, dh, danchor, eps=1e-3):
        _mat(h, "h"), _flat(g, "g", F32), _flat(out, "out", F32), _flat(anchor, "anchor", F32), _flat(W, "W", BF16)
        _flat(dd_op, "dd_op", BF16), _flat(dh, "dh", BF16), _flat(danchor, "danchor", F32)
        dz = torch.empty_like(out)
        self.box_refine_bwd(out, anchor, g, dz, danchor, eps)
        self.launches -= 1
        _store(dd_op, dz)
        _store(dh, (_f(dd_op) @ _f(W)) * (_f(h) > 0))
        self.launches += 1

    def cls_gather(self, X, video, pos, Y, qk_op, y_op, r):
        _flat(X, "X", F32), _flat(video, "video", F32), _flat(pos, "pos", F32), _flat(Y, "Y", F32)
        _flat(qk_op, "qk_op", BF16), _flat(y_op, "y_op", BF16)
        Y.copy_(torch.cat([video, X[:, r, :]], 0))
        if qk_op is not None:
            _store(qk_op, Y + pos)
        if y_op is not None:
            _store(y_op, Y)
        self.launches += 1

    def cls_scatter(self, Y, X, X_op, r, qk_next=None, pos=None):
        _flat(Y, "Y", F32), _flat(X, "X", F32), _flat(X_op, "X_op", BF16), _flat(qk_next, "qk_next", BF16), _flat(pos, "pos", F32)
        X[:, r, :] = Y[1:]
        if X_op is not None:
            X_op[:, r, :] = Y[1:].to(BF16)
        if qk_next is not None:
            qk_next[:, r, :] = (Y[1:] + pos[:, r, :]).to(BF16)
        self.launches += 1

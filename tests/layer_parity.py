"""TEST INFRASTRUCTURE: per-layer ("teacher-forced") parity of the bf16 tensor-core path against the kernel-matched oracle.

Why per layer.  The bf16 path keeps activations between kernels in bf16.  A post-norm transformer of 6 + 6
encoder and 6 + 6 decoder layers amplifies bf16 rounding noise chaotically: the ORACLE ITSELF, with identical rounding
points, moves by 3e-3 ... 1.5e-2 (outputs) and up to 0.4 (gradients, relative to max) when its accumulation dtype goes
from fp32 to fp64 (tests/test_oracle_golden.py::test_bf16_noise_floor...).  An end-to-end 1e-3 gate therefore cannot be met
by any bf16 implementation and would say nothing about kernel correctness.  What can be gated tightly is every layer on its
own: the layer's real input is taken from a full forward pass of the product path (so the statistics are the real ones),
the layer is run by the product path and by ``oracle.stcat_oracle`` with ``BF16_KERNEL`` precision (bf16 rounding at the
same points as the kernels) on that same input, and outputs / gradients are compared.  A wrong rounding point, a wrong
scale, a mis-indexed mask or a broken kernel shows up at >= 4e-3 (one bf16 ulp) on its layer; matched layers agree to
~1e-4 ... 1e-3 (isolated rounding flips caused by fp32 summation order).

Used by tests/test_gpu_bf16_parity.py (on the B200 through the C ABI) and by tests/test_bf16_parity_emu.py (CPU: the
host composition over the torch emulation of the C ABI -- checks this harness and the oracle's rounding model without a GPU).
"""
import contextlib

import torch

from oracle import stcat_oracle as O
from stcat_b200 import decoder as dec
from stcat_b200 import encoder as enc
from stcat_b200 import ops
from stcat_b200.nested import NestedTensor

NHEAD = 8


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def grad_ok(a, b, min_cos=0.999, max_l2=5e-2):
    """the gradient gate of a single layer: cosine >= 0.999 and relative L2 error <= 5e-2 against the oracle's fp64 gradients on
    bf16-rounded operands.  Max-abs is reported, not gated: the kernels round gradient operands (dy, dS, dz) to bf16, and an
    isolated bf16 flip in the forward (e.g. a ReLU unit within one ulp of zero) moves single gradient elements by several
    percent of the maximum -- the torch emulation of the kernels' rounding shows the same (tests/test_bf16_parity_emu.py)."""
    return cosine(a, b) >= min_cos and rel_l2(a, b) <= max_l2


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


@contextlib.contextmanager
def record_layers(records):
    """Record (kind, module, inputs, outputs) of every encoder / decoder layer executed inside the block."""
    e_run, b_run, t_run = enc.TransformerEncoderLayer.run, dec.TransformerDecoderLayer.run, dec.TimeDecoderLayer.run
    c_init = dec._Ctx.__init__
    td_run = dec.TimeDecoder.run
    cl = lambda t: None if t is None else t.detach().clone()

    def e_wrap(self, x, x_op, pos, key_mask, B, L, pos_cls=None, qk_op=None, pre=None):
        x_in = cl(x)  # the stream is edited in place afterwards (frame-CLS row exchange): keep copies
        out = e_run(self, x, x_op, pos, key_mask, B, L, pos_cls=pos_cls, qk_op=qk_op, pre=pre)
        records.append(("enc", self, dict(x=x_in, pos=cl(pos), key_mask=key_mask, B=B, L=L), dict(y=cl(out[0]))))
        return out

    def c_wrap(self, idx, mem, mem_pos, key_mask, n_mem_tokens, operands=None, stream=None):
        c_init(self, idx, mem, mem_pos, key_mask, n_mem_tokens, operands, stream)
        if mem is None:  # fused glue (ops.mem_operands): the fp32 memory / positions are rows 1.. of the encoder stream
            d = stream[0].shape[-1]
            mem, mem_pos = stream[0][:, 1:].reshape(-1, d), stream[1][:, 1:].reshape(-1, d)
        self._mem, self._mem_pos = cl(mem), cl(mem_pos)

    def b_wrap(self, c, tgt, tgt_op, query_pos, query_time, time_op, query_sine, sine_op, is_first, mem_kv, pos_op=None):
        ins = dict(c=c, tgt=cl(tgt), query_pos=cl(query_pos), query_time=cl(query_time), query_sine=cl(query_sine),
                   is_first=is_first)
        out = b_run(self, c, tgt, tgt_op, query_pos, query_time, time_op, query_sine, sine_op, is_first, mem_kv, pos_op=pos_op)
        records.append(("box", self, ins, dict(y=cl(out[0]))))
        return out

    def t_wrap(self, c, tgt, tgt_op, query_pos, query_pos_frames, qpos_plus_time, mem_kv):
        ins = dict(c=c, tgt=cl(tgt), query_pos=cl(query_pos), qpt=cl(qpos_plus_time))
        out = t_run(self, c, tgt, tgt_op, query_pos, query_pos_frames, qpos_plus_time, mem_kv)
        records.append(("time", self, ins, dict(y=cl(out[0]), weights=cl(out[2]))))
        return out

    def td_wrap(self, c, tgt, query_pos, query_time, mem_kv):
        c._query_time = cl(query_time)
        return td_run(self, c, tgt, query_pos, query_time, mem_kv)

    enc.TransformerEncoderLayer.run, dec.TransformerDecoderLayer.run, dec.TimeDecoderLayer.run = e_wrap, b_wrap, t_wrap
    dec._Ctx.__init__, dec.TimeDecoder.run = c_wrap, td_wrap
    try:
        yield records
    finally:
        enc.TransformerEncoderLayer.run, dec.TransformerDecoderLayer.run, dec.TimeDecoderLayer.run = e_run, b_run, t_run
        dec._Ctx.__init__, dec.TimeDecoder.run = c_init, td_run


def run_full(model, inp, device):
    mv = lambda x: x.to(device)
    videos = NestedTensor(mv(inp["vis_features"]), mv(inp["vis_mask"]), inp["durations"])
    return model(videos, mv(inp["vis_pos"]), (mv(inp["text_mask"]), mv(inp["text_memory"]), None))


def _sf(x, B, L):
    """batch-major rows [B*L, c] -> sequence-first [L, B, c] (the oracle / reference layout), on the CPU"""
    return x.detach().cpu().view(B, L, -1).transpose(0, 1).contiguous()


def _bm(x):
    """sequence-first [L, B, c] -> batch-major rows [B*L, c]"""
    return x.transpose(0, 1).reshape(-1, x.shape[-1])


def _layer_params(P, prefix, dtype, grad):
    out = {}
    for k, v in P.items():
        if k.startswith(prefix + "."):
            out[k] = v.detach().cpu().to(dtype).clone().requires_grad_(grad)
    return out


def _oracle_layer(kind, P, prefix, ins, durations, prec, from_scratch=True):
    """Runs the oracle's layer function on the recorded inputs (given as leaves in ``ins``).  Returns a dict of
    batch-major outputs."""
    if kind == "enc":
        y = O.encoder_layer(P, prefix, ins["x"], ins["mask"], ins["pos"], NHEAD, prec)
        return dict(y=_bm(y))
    if kind == "box":
        y, _ = O.box_decoder_layer(P, prefix, ins["tgt"], ins["mem"], ins["query_mask"], ins["mem_mask"], ins["pos"],
                                   ins["query_pos"], ins["query_time"], ins["query_sine"], durations, ins["is_first"], NHEAD,
                                   prec, from_scratch)
        return dict(y=_bm(y))
    y, w = O.time_decoder_layer(P, prefix, ins["tgt"], ins["mem"], ins["query_mask"], ins["mem_mask"], ins["pos"],
                                ins["query_pos"], ins["query_time"], durations, NHEAD, prec)
    return dict(y=_bm(y), weights=w)


def _oracle_inputs(kind, ins, dtype, grad):
    """recorded device tensors -> the oracle's sequence-first CPU tensors; differentiable inputs become leaves"""
    leaf = lambda t: t.to(dtype).clone().requires_grad_(grad)
    if kind == "enc":
        B, L = ins["B"], ins["L"]
        return dict(x=leaf(_sf(ins["x"], B, L)), pos=_sf(ins["pos"], B, L).to(dtype), mask=ins["key_mask"].cpu().bool())
    c = ins["c"]
    b, t, n, M = c.b, c.t, c.n, c.M
    out = dict(tgt=leaf(_sf(ins["tgt"], b, t)), query_pos=leaf(_sf(ins["query_pos"], b, t)),
               mem=leaf(_sf(c._mem, n, M)), pos=_sf(c._mem_pos, n, M).to(dtype),
               query_mask=c.query_mask.cpu().bool(), mem_mask=c.key_mask.cpu().bool())
    if kind == "box":
        out["query_time"] = _sf(ins["query_time"], b, t).to(dtype)
        out["query_sine"] = leaf(_sf(ins["query_sine"], b, t))
        out["is_first"] = ins["is_first"]
    else:
        out["query_time"] = _sf(c._query_time, b, t).to(dtype)
    return out


def check_forward(records, names, P, durations, from_scratch=True):
    """every recorded layer vs the kernel-matched oracle on the same input.  Returns {layer name: {output: rel-max err}}."""
    errs = {}
    with torch.no_grad():
        for kind, mod, ins, outs in records:
            prefix = names[id(mod)]
            oi = _oracle_inputs(kind, ins, torch.float32, False)
            ref = _oracle_layer(kind, {k: v.detach().cpu() for k, v in P.items() if k.startswith(prefix + ".")}, prefix, oi,
                                durations, O.BF16_KERNEL, from_scratch)
            errs[prefix] = {k: rel_max(outs[k], ref[k]) for k in ref}
    return errs


def _replay(kind, mod, ins, gseed):
    """Runs one layer of the product path again on its recorded input with gradients enabled; returns
    (upstream gradient g, {input name: grad}, {param name (relative): grad})."""
    for p in mod.parameters():
        p.grad = None
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(gseed)
    if kind == "enc":
        x = leaf(ins["x"])
        y, _ = mod.run(x, None, ins["pos"], ins["key_mask"], ins["B"], ins["L"])
        leaves = dict(x=x)
    else:
        c0 = ins["c"]
        mem = leaf(c0._mem)
        c = dec._Ctx(c0.idx, mem, c0._mem_pos, c0.key_mask, c0.M)
        tgt, qpos = leaf(ins["tgt"]), leaf(ins["query_pos"])
        leaves = dict(tgt=tgt, query_pos=qpos, mem=mem)
        if kind == "box":
            qsine = leaf(ins["query_sine"])
            leaves["query_sine"] = qsine
            y, _ = mod.run(c, tgt, None, qpos, ins["query_time"], None, qsine, None, ins["is_first"],
                           mod.memory_side(c, ins["is_first"]))
        else:
            y, _, _ = mod.run(c, tgt, None, qpos, c.frames(qpos), qpos + c0._query_time, mod.memory_side(c))
    g = torch.randn(y.shape, generator=gen)
    y.backward(g.to(y.device).clone())  # PutRowsFn edits its incoming gradient in place
    return g, {k: v.grad for k, v in leaves.items()}, {k: p.grad for k, p in mod.named_parameters() if p.grad is not None}


def check_backward(records, names, P, durations, select=None, from_scratch=True, gseed=11):
    """Gradients of single layers (inputs and parameters) for a random upstream gradient: product path vs the oracle in fp64
    on bf16-rounded operands / stores.  Returns {layer: {tensor: (rel-max err, cosine, grad_ok, rel-L2 err)}}."""
    prec = O.Prec(torch.float64, "bf16", True) if ops.get_precision() == "bf16" else O.Prec(torch.float64)
    res = {}
    for kind, mod, ins, outs in records:
        prefix = names[id(mod)]
        if select is not None and not select(prefix):
            continue
        g, gin, gpar = _replay(kind, mod, ins, gseed)
        oi = _oracle_inputs(kind, ins, torch.float64, True)
        Pl = _layer_params(P, prefix, torch.float64, True)
        ref = _oracle_layer(kind, Pl, prefix, oi, durations, prec, from_scratch)
        if kind == "enc":
            B, L = ins["B"], ins["L"]
        else:
            B, L = ins["c"].b, ins["c"].t
        ref["y"].backward(g.double().view(B, L, -1).reshape(B * L, -1))
        r = {}
        for k, gv in gin.items():
            ok = {"x": "x", "tgt": "tgt", "query_pos": "query_pos", "query_sine": "query_sine", "mem": "mem"}[k]
            gref = oi[ok].grad
            if gref is None:
                continue
            gref_bm = _bm(gref)
            r["d" + k] = (rel_max(gv, gref_bm), cosine(gv, gref_bm), grad_ok(gv, gref_bm), rel_l2(gv, gref_bm))
        for k, gv in gpar.items():
            gref = Pl[f"{prefix}.{k}"].grad
            if gref is None or float(gref.abs().max()) < 1e-9:
                continue  # e.g. key-projection biases: a per-row constant of the scores has no gradient through the softmax
            r[k] = (rel_max(gv, gref), cosine(gv, gref), grad_ok(gv, gref), rel_l2(gv, gref))
        res[prefix] = r
    return res


def module_names(model):
    return {id(m): n for n, m in model.named_modules()}

"""Data-parallel plumbing of the hot path: one clip per GPU, one process per GPU (the reference's only regime,
datasets/build.py:151,157), gradients summed across ranks once per step.

The reference wraps the model in ``DistributedDataParallel(find_unused_parameters=True)`` (train_net.py:31-36):
bucketed NCCL all-reduces plus a per-iteration unused-parameter bitmap all-reduce, needed because
``ground_encoder.fusion.*`` and the six ``ca_qtime_proj.*`` never receive a gradient.  Here
every hot-path gradient is a view of ONE contiguous fp32 buffer: the wgrad kernels accumulate straight into it
(``ops.set_grad_fusion``), one memset clears it, ONE all-reduce per step exchanges it, and parameters the
forward never touches simply keep their zero slice (no unused-parameter search).
"""
from __future__ import annotations

import torch


class FlatGrads:
    """All gradients of ``model`` as views of one contiguous fp32 buffer.

    ``groups`` (optional): ordered ``[(name, [params...]), ...]``; each group occupies one contiguous range
    (``self.ranges[name] = (lo, hi)``) so it can be all-reduced on its own as soon as its gradients are complete
    (``GradSync``).  Parameters not named by any group go to a trailing group ``"rest"``."""

    ALIGN = 64  # floats

    def __init__(self, model, groups=None):
        seen, ordered, self.ranges = set(), [], {}
        spans = []
        for name, ps in (groups or []):
            lo = len(ordered)
            for p in ps:
                if id(p) not in seen:
                    seen.add(id(p))
                    ordered.append(p)
            spans.append((name, lo, len(ordered)))
        lo = len(ordered)
        for p in model.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                ordered.append(p)
        spans.append(("rest", lo, len(ordered)))
        params = ordered
        # every parameter's slice starts on a 256-byte boundary: the wgrad kernels accumulate into these views with
        # TMA reduce-add (16-byte base alignment required; an unaligned view would silently take the slow SIMT path)
        A = self.ALIGN
        starts, o = [], 0
        for p in params:
            starts.append(o)
            o += (p.numel() + A - 1) // A * A
        n = o
        self.buf = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        self.params = params
        self.starts = list(starts)  # offset of every parameter of ``params`` in the flat buffer (optim.FusedAdamW)
        for p, st in zip(params, starts):
            p.grad = self.buf[st:st + p.numel()].view_as(p)
        starts.append(n)
        for name, a, b in spans:
            if b > a:
                self.ranges[name] = (starts[a], starts[b])
        self.numel = n

    def zero(self):
        from . import ops

        ops.join_leaf_streams()  # leaf-stream weight gradients of the previous step must have landed
        self.buf.zero_()

    def zero_async(self):
        """``zero()`` on a side stream forked from the current one: the fill (180 MB, ~25 us) runs under the forward pass, which
        never touches the gradient buffer.  ``wait_zero()`` must be called on the stream that starts the backward pass."""
        from . import ops

        ops.join_leaf_streams()
        if not self.buf.is_cuda:
            self.buf.zero_()
            return
        cur = torch.cuda.current_stream()
        side = getattr(self, "_zero_stream", None)
        if side is None:
            side = self._zero_stream = torch.cuda.Stream(self.buf.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.buf.zero_()
            self._zero_event = side.record_event()

    def wait_zero(self):
        ev = getattr(self, "_zero_event", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._zero_event = None

    def all_reduce(self, average: bool = False):
        """Sum (or mean) of the flat buffer over the default process group; no-op without one."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        dist.all_reduce(self.buf)
        if average:
            self.buf.div_(dist.get_world_size())


def hot_path_groups(model):
    """Gradient groups of ``STCATHotPath`` in the order their gradients complete during backward: decoder + heads
    first (everything downstream of the encoder output), then the encoder blocks from the last to the first."""
    enc = model.ground_encoder.encoder
    dec = [p for m in (model.ground_decoder, model.bbox_embed, model.temp_embed, getattr(model, "action_embed", None))
           if m is not None for p in m.parameters()]
    groups = [("decoder", dec)]
    for i in reversed(range(enc.num_layers)):
        groups.append((f"enc{i}", list(enc.spatial_layers[i].parameters()) + list(enc.temporal_layers[i].parameters())))
    return groups


class GradSync:
    """All-reduces ranges of the flat gradient buffer on a side stream while the backward pass is still running.

    The hot path's backward runs decoder -> encoder block 5 -> ... -> block 0, and each block's weight gradients are
    final when the gradient of that block's INPUT has been produced.  ``attach(out, model)`` registers autograd hooks
    on exactly those tensors (the encoder outputs handed to the decoder, and each block's input), so each range's
    NCCL all-reduce is enqueued the moment it is complete and overlaps the rest of the backward; ``finish()`` reduces
    what is left and joins the side stream.  Works eagerly and under CUDA-graph capture (the collectives are captured
    on the side stream like any other kernel).  Without a process group every call is a no-op."""

    def __init__(self, flat: FlatGrads, device=None, average: bool = True):
        """``average`` (default): gradients are AVERAGED over the ranks, like the torch DistributedDataParallel wrapper
        this replaces (train_net.py:31-36); the loss already divides by num_boxes / world_size (criterion.py:177-179),
        which assumes that averaging.  ``average=False`` leaves the sum (then pass ``grad_scale=1/world`` to FusedAdamW)."""
        import torch.distributed as dist

        self.flat = flat
        self.average = bool(average)
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.stream = torch.cuda.Stream(device) if (self.active and flat.buf.is_cuda) else None
        self.done = set()
        self.extra_streams = ()
        # While an all-reduce is in flight the persistent kernels of the backward pass leave ``nccl_ctas`` SMs to it
        # (include/stcat_b200.h stcat_set_sm_cap): NCCL's CTAs stay resident for the whole collective, and a persistent GEMM /
        # attention CTA that cannot become resident next to one would hold its share of the tiles until the collective ends.
        # Keep NCCL_MAX_CTAS (read by NCCL when the communicator is created; bench.py sets it) at this value.  0 = no cap.
        import os

        self.nccl_ctas = int(os.environ.get("STCAT_NCCL_CTAS", os.environ.get("NCCL_MAX_CTAS", "24") or 0))
        self._capped = False

    def prepare(self, model):
        """Observe the encoder's block inputs (call once, before the first forward)."""
        if self.active:
            model.ground_encoder.encoder.layer_input_callback = self._on_layer_input

    def _on_layer_input(self, i, x):
        # gradient of block i's input ready  =>  blocks >= i (spatial and temporal) have accumulated their gradients
        if x.requires_grad and torch.is_grad_enabled():
            x.register_hook(lambda g, i=i: self._reduce(f"enc{i}"))

    def begin_step(self):
        self.done = set()

    def _reduce(self, name):
        import torch.distributed as dist

        if not self.active or name in self.done or name not in self.flat.ranges:
            return
        self.done.add(name)
        lo, hi = self.flat.ranges[name]
        chunk = self.flat.buf[lo:hi]
        if self.stream is None:
            dist.all_reduce(chunk)  # gloo (CPU tests): no AVG op
            if self.average:
                chunk.div_(dist.get_world_size())
            return
        from . import ops

        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        for st in (*self.extra_streams, *ops.leaf_streams()):  # gradient kernels of this range may have been enqueued on these too
            self.stream.wait_stream(st)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(chunk, op=dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM)
        if self.nccl_ctas > 0 and not self._capped:
            sms = torch.cuda.get_device_properties(chunk.device).multi_processor_count
            ops.get_backend().set_sm_cap(max(sms - self.nccl_ctas, sms // 2))
            self._capped = True

    def attach(self, out: dict):
        """Hook for one forward pass (``out`` = STCATHotPath's output dict): the decoder + heads range is complete
        when the gradients of all tensors the encoder handed to the decoder have been produced.  Plain tensor hooks
        with a counter (torch's register_multi_grad_hook builds reference cycles that keep the step's autograd
        graph -- and its AccumulateGrad nodes, bound to the stream of that step -- alive into the next step, which
        breaks CUDA-graph capture)."""
        if not self.active:
            return
        mc = out["_memory_cache"]
        if mc.get("_stream_used"):
            # the decoder read the encoder's frame-major stream itself (ops.mem_operands): the views below receive no gradient
            boundary = [t for t in (mc["_stream"][0], mc["videos_cls"]) if t.requires_grad]
        else:
            boundary = [t for t in (mc["encoded_memory"], mc["frames_cls"], mc["videos_cls"]) if t.requires_grad]
        state = {"seen": 0, "need": len(boundary)}

        def hook(g):
            state["seen"] += 1
            if state["seen"] == state["need"]:
                self._reduce("decoder")

        for t in boundary:
            t.register_hook(hook)

    def finish(self):
        from . import ops

        ops.join_leaf_streams()
        for name in list(self.flat.ranges):
            self._reduce(name)
        if self._capped:
            ops.get_backend().set_sm_cap(0)
            self._capped = False
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

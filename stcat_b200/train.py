"""A training step of the hot path as one replayable CUDA graph (what ``bench.py`` times, usable from the reference's loop).

The reference's iteration (scripts/train_net.py:97-143) is: forward, loss, ``optimizer.zero_grad``, backward, clip, step, EMA
-- issued kernel by kernel from Python.  On a B200 the hot path's ~1100 C-ABI calls per step take ~6 ms of device time but
~50 ms to ISSUE from Python (profiles/r2_b_bench_eager.json), so the eager loop is host-bound 8x over.  ``GraphedStep`` captures
the whole step once (all side streams, the NCCL gradient all-reduce of ``dp.GradSync`` and the fused optimizer included) and
replays it per iteration; new inputs are copied into the captured step's static input tensors.

Train-mode dropout under replay: kernel arguments are frozen at capture, so the masks' (seed, offset) are too.  The step owns
a device-resident counter (``stcat_set_dropout_step``) that every dropout kernel folds into its seed when it runs, and the
captured step increments it after the backward pass: every replay draws fresh masks, forward and backward of one step agree.

    step = GraphedStep(lambda vis, pos, txt: fwd_loss_bwd_opt(vis, pos, txt), {"vis": vis0, "pos": pos0, "txt": txt0})
    for batch in loader:
        loss = step(vis=batch_vis, pos=batch_pos, txt=batch_txt)   # device tensor; .item() only when logging
"""
from __future__ import annotations

from typing import Callable, Dict

import torch

from . import ops


class GraphedStep:
    """``fn(**static_inputs) -> scalar loss tensor`` must do the whole iteration on the given tensors (zero the gradients,
    forward, loss, backward, gradient sync, optimizer step).  Shapes are fixed at construction: one graph per clip shape
    (T, H, W, L); build one instance per shape bucket.  ``use_graph=False`` runs the same step eagerly (debugging)."""

    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Dict[str, torch.Tensor], use_graph: bool = True,
                 warmup: int = 3, dropout_counter: bool = True):
        dev = next(iter(example_inputs.values())).device
        self.fn = fn
        self.static = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in example_inputs.items()}
        self.loss = torch.zeros((), device=dev)
        self.graph = None
        self.counter = None
        if dropout_counter and dev.type == "cuda":
            self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
            ops.get_backend().set_dropout_step(self.counter)
        for _ in range(max(1, warmup)):  # eager: fills the bf16 weight cache / index caches, lets NCCL open its channels
            self._eager()
        if use_graph and dev.type == "cuda":
            torch.cuda.synchronize(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            # The step's main stream is a high-priority stream: the dependent chain of the backward pass (data gradients,
            # attention, LayerNorm) is scheduled ahead of the weight gradients on the leaf streams (ops.set_leaf_streams) and of
            # the collectives, which only have to be done by the end of the step.
            # thread_local: the NCCL watchdog thread polls events while the collectives are being captured
            import os

            main = torch.cuda.Stream(dev, priority=-1 if os.environ.get("STCAT_CHAIN_PRIO", "1") != "0" else 0)
            with torch.cuda.graph(g, stream=main, capture_error_mode="thread_local"):
                self._eager()
            torch.cuda.synchronize(dev)
            self.graph = g

    def _eager(self):
        total = self.fn(**self.static)
        self.loss.copy_(total.detach())
        if self.counter is not None:
            self.counter.add_(1)

    def replay(self) -> torch.Tensor:
        """one step on whatever the static inputs currently hold"""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager()
        return self.loss

    @torch.no_grad()
    def load_inputs(self, **inputs):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)

    def __call__(self, **inputs) -> torch.Tensor:
        self.load_inputs(**inputs)
        return self.replay()

    def close(self):
        if self.counter is not None:
            ops.get_backend().set_dropout_step(None)
            self.counter = None

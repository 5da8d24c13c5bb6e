"""Evaluation path around the hot path: reference engine/evaluate.py:20-129.

The reference evaluates a clip in TWO forward passes -- its even and its odd frames (``videos.subsample(2, 0 / 1)``,
evaluate.py:97-110) -- and merges them: boxes of both passes, linearly interpolated over the frame ids in between
(``linear_interp``), and the union of the two predicted temporal segments.  Here

* the two passes are ONE forward over a ragged batch of 2b videos (the even and the odd frames of every clip are two
  "videos" with durations ceil(T/2) and floor(T/2); the hot path's masked / ragged general path makes them independent, so the
  result equals the two separate passes);
* ``PostProcess`` (T x T start/end scoring, ``stcat_sted_score``) and the box interpolation (``stcat_box_interp``) run on
  the device; the only host transfers are the final boxes / segment indices.

``double_pass`` takes the inputs at the hot path's seam (``input_proj``-ed visual features, the positional encoding, the text
encoder's output tuple), so it serves both ``STCATHotPath`` and a full model that calls it after its backbone.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import ops
from .nested import NestedTensor
from .pipeline import PostProcess


def even_odd_batch(videos: NestedTensor, vis_pos: torch.Tensor, texts):
    """[clip 0 even, clip 0 odd, clip 1 even, ...] as one ragged batch; text rows repeated per pass."""
    durations = list(videos.durations)
    t_all, m_all, p_all, durs = [], [], [], []
    for v, m, p in zip(torch.split(videos.tensors, durations, 0), torch.split(videos.mask, durations, 0),
                       torch.split(vis_pos, durations, 0)):
        for start in (0, 1):
            t_all.append(v[start::2]); m_all.append(m[start::2]); p_all.append(p[start::2])
            durs.append(t_all[-1].shape[0])
    text_mask, text_memory, tok = texts
    texts2 = (text_mask.repeat_interleave(2, 0), text_memory.repeat_interleave(2, 1), tok)
    return NestedTensor(torch.cat(t_all, 0), torch.cat(m_all, 0), durs), torch.cat(p_all, 0), texts2


@torch.no_grad()
def double_pass(hot_path, videos: NestedTensor, vis_pos: torch.Tensor, texts, targets: Sequence[dict],
                postprocessor: PostProcess = None):
    """evaluate.py:81-124 for one batch.  ``targets[i]``: {"item_id", "ori_size" (h, w), "frame_ids" (list[int], one per
    frame of clip i), optional "qtype"}.  Returns (bbox_pred {item_id: {frame_id: [[x1, y1, x2, y2]]}}, temp_pred
    {item_id: {"sted": [start, end] (, "qtype")}}) like the reference's ``do_eval`` hands to its evaluator."""
    post = postprocessor or PostProcess()
    b = len(targets)
    if any(d < 2 for d in videos.durations):
        raise ValueError("the even/odd evaluation needs at least 2 frames per clip")
    v2, p2, t2 = even_odd_batch(videos, vis_pos, texts)
    out = hot_path(v2, p2, t2)
    durs = v2.durations
    t = max(durs)
    dev = out["pred_boxes"].device
    sizes = torch.tensor([list(tg["ori_size"]) for tg in targets for _ in range(2) for _ in range(t)], device=dev, dtype=torch.float32)
    frame_ids2 = [list(tg["frame_ids"])[s::2] for tg in targets for s in (0, 1)]
    boxes, steds = post(out, sizes, frame_ids2, durs)
    boxes = boxes.view(2 * b, t, 4)
    be = ops.get_backend()
    bbox_pred: Dict = {}
    temp_pred: Dict = {}
    for i, tg in enumerate(targets):
        fids = list(tg["frame_ids"])
        de, do = durs[2 * i], durs[2 * i + 1]
        # even / odd predictions back in frame order, then every frame id in between by linear interpolation
        merged = torch.empty(len(fids), 4, device=dev, dtype=torch.float32)
        merged[0::2] = boxes[2 * i, :de]
        merged[1::2] = boxes[2 * i + 1, :do]
        order = sorted(range(len(fids)), key=lambda j: fids[j])
        ids = torch.tensor([fids[j] for j in order], device=dev, dtype=torch.int64)
        src = merged.index_select(0, torch.tensor(order, device=dev)).contiguous()
        first, last = int(fids[order[0]]), int(fids[order[-1]])
        dense = torch.empty(last - first + 1, 4, device=dev, dtype=torch.float32)
        be.box_interp(ids, src, dense, first)
        rows = dense.cpu().tolist()
        bbox_pred[tg["item_id"]] = {first + j: [rows[j]] for j in range(len(rows))}
        s1, s2 = steds[2 * i], steds[2 * i + 1]
        temp_pred[tg["item_id"]] = {"sted": [min(s1[0], s2[0]), max(s1[1], s2[1])]}
        if "qtype" in tg:
            temp_pred[tg["item_id"]]["qtype"] = tg["qtype"]
    return bbox_pred, temp_pred

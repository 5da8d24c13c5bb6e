"""Text side of ``STCATNet`` in front of the hot path: reference models/language_model/bert.py:40-110.

* ``FeatureResizer``: ``Linear(768, 256)`` -> ``LayerNorm(eps 1e-12)`` -> dropout, over this package's C ABI (one GEMM + one
  LayerNorm kernel for the whole [L * b (+ b pooled rows), 768] matrix: the reference runs the resizer twice, on the sequence
  and on the pooled vector).  Same ``state_dict`` entries (``fc.weight``, ``fc.bias``, ``layer_norm.weight``, ``layer_norm.bias``).
* ``TextEncoder``: ``Roberta``-shaped wrapper.  The language model itself is the Hugging Face ``RobertaModel`` (library code, like
  the reference); what changes is around it: the tokenizer call is hoisted out of ``forward`` (``tokenize`` once per
  caption on the data-loader side; ``forward`` also accepts the token tensors), and the output tuple is the one the hot path
  takes: ``(text_attention_mask [b, L] bool True = pad, text_memory [L, b, 256], tokenized), text_cls [b, 256]``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from .params import LinearP, NormP


class FeatureResizer(nn.Module):
    """bert.py:91-110."""

    def __init__(self, input_feat_size: int, output_feat_size: int, dropout: float, do_ln: bool = True):
        super().__init__()
        self.do_ln = do_ln
        self.fc = LinearP(input_feat_size, output_feat_size)
        self.layer_norm = NormP(output_feat_size, eps=1e-12)
        self.eps = 1e-12
        self.dropout_p = float(dropout)

    def forward(self, encoder_features: torch.Tensor) -> torch.Tensor:
        shape = encoder_features.shape
        x = ops.linear(encoder_features.reshape(-1, shape[-1]), self.fc.weight, self.fc.bias)
        if self.do_ln:
            x = ops.layer_norm(x, None, self.layer_norm.weight, self.layer_norm.bias, eps=self.eps)
        x = ops.dropout(x, self.dropout_p if self.training else 0.0)
        return x.view(*shape[:-1], -1)


class TextEncoder(nn.Module):
    """bert.py:40-89 ``Roberta(name, outdim, freeze)``.  ``body`` / ``tokenizer`` may be passed in (offline construction:
    ``RobertaModel(RobertaConfig(...))``); otherwise they are loaded by name like the reference."""

    def __init__(self, name: str = "roberta-base", outdim: int = 256, freeze: bool = False, body: Optional[nn.Module] = None,
                 tokenizer=None):
        super().__init__()
        if body is None:
            from transformers import RobertaModel, RobertaTokenizerFast

            body = RobertaModel.from_pretrained(name)
            tokenizer = tokenizer or RobertaTokenizerFast.from_pretrained(name)
        self.body = body
        self.tokenizer = tokenizer
        if freeze:
            for p in self.body.parameters():
                p.requires_grad_(False)
        self.resizer = FeatureResizer(self.body.config.hidden_size, outdim, dropout=0.1)

    def tokenize(self, texts, device=None):
        """The tokenizer call of bert.py:60-61, to be made once per caption outside the step (collate function)."""
        tok = self.tokenizer.batch_encode_plus(texts, padding="longest", return_tensors="pt")
        return tok if device is None else tok.to(device)

    def forward(self, texts, device=None):
        """texts: list[str] (tokenised here, like the reference) or the output of ``tokenize``."""
        tokenized = self.tokenize(texts, device) if isinstance(texts, (list, tuple)) else texts
        enc = self.body(input_ids=tokenized["input_ids"], attention_mask=tokenized["attention_mask"])
        seq = enc.last_hidden_state  # [b, L, 768]
        b, L, c = seq.shape
        # one resizer pass over the L * b sequence rows (sequence-first, as the hot path wants them) and the b pooled rows
        rows = torch.cat([seq.transpose(0, 1).reshape(L * b, c), enc.pooler_output], 0)
        out = self.resizer(rows)
        text_memory = out[:L * b].view(L, b, -1)
        text_cls = out[L * b:]
        mask = tokenized["attention_mask"].ne(1).bool()
        return (mask, text_memory, tokenized), text_cls

"""The drop-in claim, end to end, on CPU: the UNMODIFIED reference ``STCATNet`` (models/pipeline.py) is built twice -- once
stock, once with ``build_encoder`` / ``build_decoder`` replaced by ``stcat_b200``'s (the 2-line change of INTEGRATION.md
section 1) -- the second model loads the first one's ``state_dict`` strictly, and both are run on the same ragged two-video
batch.  Everything around the two factories is the reference's own code: ``STCATNet.__init__`` / ``forward``, its
``NestedTensor``, ``input_proj``, the head MLPs (assigned into our decoder, pipeline.py:50), ``inverse_sigmoid``.

The vision backbone and the text encoder sit upstream of the hot path and need pretrained downloads (ResNet-101, RoBERTa),
so they are replaced -- for BOTH models -- by small stand-ins with the same interfaces.  The arithmetic of ``stcat_b200``
runs through the torch emulation of its C ABI (no GPU in the build container); the kernels themselves are covered by the
``-m gpu`` tests.  Needs /root/reference: skipped elsewhere (the GPU box)."""
import pytest
import torch
from torch import nn

from helpers import rel_err
from oracle.ref_import import import_reference, reference_available
from stcat_b200 import ops

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the reference checkout under /root/reference")


class TinyBackbone(nn.Module):
    """stand-in for models/vision_model (Joiner(Backbone, PositionEmbeddingSine)): NestedTensor of frames ->
    (NestedTensor of stride-32 features, sine position embedding), ``num_channels``"""

    def __init__(self, ref, hidden):
        super().__init__()
        self.num_channels = 48
        self.conv = nn.Conv2d(3, self.num_channels, 32, stride=32)
        self.pos = ref.PositionEmbeddingSine(hidden // 2, normalize=True)
        self.ref = ref

    def forward(self, videos):
        x = self.conv(videos.tensors)
        m = torch.nn.functional.interpolate(videos.mask[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
        out = self.ref.NestedTensor(x, m, videos.durations)
        return out, self.pos(out).to(x.dtype)


class TinyText(nn.Module):
    """stand-in for models/language_model.Roberta: (texts, device) -> ((pad mask [b, L], memory [L, b, d], tokenized), cls)"""

    def __init__(self, hidden):
        super().__init__()
        self.emb = nn.Embedding(97, hidden)

    def forward(self, texts, device):
        L = max(len(t.split()) for t in texts) + 2
        ids = torch.zeros(len(texts), L, dtype=torch.long)
        mask = torch.ones(len(texts), L, dtype=torch.bool)
        for i, t in enumerate(texts):
            w = [1] + [3 + sum(map(ord, tok)) % 90 for tok in t.split()] + [2]
            ids[i, :len(w)] = torch.tensor(w)
            mask[i, :len(w)] = False
        mem = self.emb(ids.to(device)).transpose(0, 1)
        return (mask.to(device), mem, None), mem[0]


@pytest.mark.parametrize("from_scratch", [True, False])
def test_reference_stcatnet_with_swapped_factories(monkeypatch, from_scratch):
    from emu_backend import EmuBackend
    import stcat_b200

    ref = import_reference()
    import models.pipeline as rp  # the reference's own module

    cfg = ref.cfg.clone()
    cfg.merge_from_list(["MODEL.STCAT.DROPOUT", 0.0, "MODEL.STCAT.FROM_SCRATCH", from_scratch])
    hidden = cfg.MODEL.STCAT.HIDDEN
    monkeypatch.setattr(rp, "build_vis_encoder", lambda c: TinyBackbone(ref, hidden))
    monkeypatch.setattr(rp, "build_text_encoder", lambda c: TinyText(hidden))
    torch.manual_seed(0)
    stock = rp.STCATNet(cfg).eval()

    monkeypatch.setattr(rp, "build_encoder", stcat_b200.build_encoder)
    monkeypatch.setattr(rp, "build_decoder", stcat_b200.build_decoder)
    torch.manual_seed(1)
    ours = rp.STCATNet(cfg).eval()
    assert type(ours.ground_encoder).__module__.startswith("stcat_b200") and type(ours.ground_decoder).__module__.startswith("stcat_b200")
    assert ours.ground_decoder.decoder.bbox_embed is ours.bbox_embed  # pipeline.py:50 works on our decoder
    sd = stock.state_dict()
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    ours.load_state_dict(sd)  # strict

    g = torch.Generator().manual_seed(3)
    durations = [3, 2]
    frames = torch.randn(sum(durations), 3, 96, 128, generator=g)
    pad = torch.zeros(sum(durations), 96, 128, dtype=torch.bool)
    pad[3:, :, 96:] = True  # the second video is narrower: padded columns, like NestedTensor.from_tensor_list produces
    texts = ["a man in red walks to the door", "the dog jumps"]
    mk = lambda: ref.NestedTensor(frames.clone(), pad.clone(), list(durations))

    def scalar(o):  # touches every output of the model, aux layers included
        s = (o["pred_boxes"] ** 2).sum() + o["pred_sted"].sum() * 0.1 + o["pred_actioness"].sum() * 0.1 + (o["weights"] ** 2).sum()
        for a in o["aux_outputs"]:
            s = s + (a["pred_boxes"] ** 2).sum() + a["pred_sted"].sum() * 0.1 + (a["weights"] ** 2).sum()
        return s

    want = stock(mk(), texts)
    scalar(want).backward()
    ops.set_backend(EmuBackend())
    try:
        got = ours(mk(), texts)
        scalar(got).backward()
    finally:
        ops.set_backend(None)
    # gradients: upstream of the hot path (they cross our backward chains), the heads, and hot-path parameters themselves
    gs, go = dict(stock.named_parameters(remove_duplicate=False)), dict(ours.named_parameters(remove_duplicate=False))
    for k in ("input_proj.weight", "text_encoder.emb.weight", "vis_encoder.conv.weight", "bbox_embed.layers.0.weight",
              "ground_encoder.encoder.spatial_layers.0.self_attn.in_proj_weight", "ground_encoder.encoder.temporal_layers.5.linear2.weight",
              "ground_decoder.decoder.layers.2.ca_kpos_proj.weight", "ground_decoder.temp_decoder.layers.4.cross_attn_image.in_proj_weight",
              "ground_decoder.template_generator.anchor_proj.weight"):
        assert rel_err(go[k].grad, gs[k].grad) < 5e-3, k
    for k, p in gs.items():  # parameters the reference never uses stay without gradient here too
        if p.grad is None:
            assert go[k].grad is None or float(go[k].grad.abs().max()) == 0.0, k
    assert set(got) == set(want)
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        assert got[k].shape == want[k].shape, k
        assert rel_err(got[k], want[k]) < 1e-4, k
    assert len(got["aux_outputs"]) == len(want["aux_outputs"]) == cfg.MODEL.STCAT.DEC_LAYERS - 1
    for a, b in zip(got["aux_outputs"], want["aux_outputs"]):
        assert set(a) == set(b)
        for k in b:
            assert rel_err(a[k], b[k]) < 1e-4, k


def test_reference_training_utilities_accept_the_swapped_model(monkeypatch):
    """train_net.py:22-36 around the model: deepcopy for the EMA copy, engine/optimizer.py's LR groups (selected by parameter
    name), update_ema over the state_dict, a checkpoint round trip through state_dict."""
    import copy
    import logging

    import stcat_b200

    ref = import_reference()
    import models.pipeline as rp
    from engine.optimizer import make_optimizer, update_ema

    cfg = ref.cfg.clone()
    hidden = cfg.MODEL.STCAT.HIDDEN
    monkeypatch.setattr(rp, "build_vis_encoder", lambda c: TinyBackbone(ref, hidden))
    monkeypatch.setattr(rp, "build_text_encoder", lambda c: TinyText(hidden))
    torch.manual_seed(0)
    stock = rp.STCATNet(cfg)
    monkeypatch.setattr(rp, "build_encoder", stcat_b200.build_encoder)
    monkeypatch.setattr(rp, "build_decoder", stcat_b200.build_decoder)
    ours = rp.STCATNet(cfg)
    ema = copy.deepcopy(ours)  # train_net.py:28
    assert ema.ground_decoder.decoder.bbox_embed is ema.bbox_embed  # the alias survives the copy, as in the reference
    groups_s = [sorted(id(p) for p in g["params"]) for g in make_optimizer(cfg, stock, logging.getLogger()).param_groups]
    names_s = {id(p): n for n, p in stock.named_parameters()}
    names_o = {id(p): n for n, p in ours.named_parameters()}
    groups_o = make_optimizer(cfg, ours, logging.getLogger()).param_groups
    for gs, go in zip(groups_s, groups_o):  # same parameter names in the same LR group
        assert sorted(names_s[i] for i in gs) == sorted(names_o[id(p)] for p in go["params"])
    with torch.no_grad():
        for p in ours.parameters():
            p.add_(1.0)
    update_ema(ours, ema, 0.9)  # engine/optimizer.py:5-22
    k = "ground_encoder.encoder.spatial_layers.3.linear1.weight"
    assert torch.allclose(ema.state_dict()[k], ours.state_dict()[k] - 0.9, atol=1e-6)
    fresh = rp.STCATNet(cfg)
    fresh.load_state_dict(copy.deepcopy(ours.state_dict()))
    assert torch.equal(fresh.state_dict()[k], ours.state_dict()[k])

"""CPU: the components either side of the hot path -- input_proj, image positional encoding,
folded-FrozenBN ResNet trunk, FeatureResizer / text encoder wrapper, the even/odd evaluation pass -- against torch / the oracle's
restatements of the reference, through the torch emulation of the C ABI."""
import math

import pytest
import torch

from oracle import stcat_oracle as O
from emu_backend import EmuBackend
from helpers import cfg_for, rel_err
from stcat_b200 import ops
from stcat_b200.nested import NestedTensor


@pytest.fixture(autouse=True)
def emu():
    ops.set_backend(EmuBackend())
    ops.set_precision("fp32")
    yield
    ops.set_backend(None)
    ops.set_precision("fp32")


def test_pos_sine_matches_reference_formula():
    from stcat_b200.vision import PositionEmbeddingSine

    mask = torch.zeros(3, 7, 9, dtype=torch.bool)
    mask[1, 5:, :] = True
    mask[2, :, 6:] = True
    mask[2, 4:, :] = True
    pe = PositionEmbeddingSine(128)
    pos = pe(NestedTensor(torch.zeros(3, 256, 7, 9), mask, [3]))
    ref = O.image_sine_pos(mask)
    assert pos.shape == ref.shape == (3, 256, 7, 9)
    assert rel_err(pos, ref) < 1e-6
    # unpadded clips: one cached frame table, expanded
    m0 = torch.zeros(4, 5, 5, dtype=torch.bool)
    a = pe(NestedTensor(torch.zeros(4, 256, 5, 5), m0, [4]))
    b = pe(NestedTensor(torch.zeros(4, 256, 5, 5), m0, [4]))
    assert a.data_ptr() == b.data_ptr() and rel_err(a, O.image_sine_pos(m0)) < 1e-6


@pytest.mark.parametrize("channels_last", [False, True])
def test_input_proj_matches_conv2d(channels_last):
    from stcat_b200.vision import InputProj

    torch.manual_seed(0)
    proj = InputProj(96, 256)
    conv = torch.nn.Conv2d(96, 256, 1)
    conv.load_state_dict(proj.state_dict())  # same names / shapes as the reference's nn.Conv2d
    x = torch.randn(5, 96, 4, 6)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = proj(xa), conv(xb)
    assert ya.shape == yb.shape and rel_err(ya, yb) < 1e-5
    assert rel_err(ya, O.input_proj({"p.weight": proj.weight, "p.bias": proj.bias}, "p", x)) < 1e-5
    assert ya.flatten(2).transpose(1, 2).is_contiguous()  # token-major rows: what the encoder's assembly reads
    g = torch.randn_like(yb)
    ya.backward(g); yb.backward(g)
    assert rel_err(xa.grad, xb.grad) < 1e-5
    assert rel_err(proj.weight.grad, conv.weight.grad) < 1e-5 and rel_err(proj.bias.grad, conv.bias.grad) < 1e-5


def test_folded_frozen_bn_trunk_matches_torchvision():
    import torchvision
    from torchvision.models._utils import IntermediateLayerGetter
    from torchvision.ops.misc import FrozenBatchNorm2d as TVFrozenBN
    from stcat_b200.vision import VisionEncoder

    torch.manual_seed(1)
    enc = VisionEncoder(name="resnet50", train_backbone=True)
    ref = IntermediateLayerGetter(torchvision.models.resnet50(weights=None, norm_layer=lambda n: TVFrozenBN(n, eps=1e-5)), {"layer4": "0"})
    sd = enc[0].body.state_dict()
    for k, v in sd.items():  # non-trivial BN statistics
        if k.endswith("running_var"):
            v.uniform_(0.5, 2.0)
        elif k.endswith("running_mean") or (k.endswith("bias") and "bn" in k):
            v.normal_(0, 0.2)
        elif k.endswith("weight") and ("bn" in k or "downsample.1" in k):
            v.uniform_(0.5, 1.5)
    ref.load_state_dict(sd)  # identical key set: what a reference checkpoint's vis_encoder.0.body.* holds
    assert set(enc.state_dict()) == {"0.body." + k for k in ref.state_dict()}
    frames = torch.randn(2, 3, 64, 96)
    mask = torch.zeros(2, 64, 96, dtype=torch.bool)
    mask[1, :, 64:] = True
    out, pos = enc(NestedTensor(frames, mask, [2]))
    y_ref = ref(frames)["0"]
    assert out.tensors.shape == y_ref.shape == (2, 2048, 2, 3)
    assert rel_err(out.tensors, y_ref) < 1e-4
    assert out.mask.shape == (2, 2, 3) and bool(out.mask[1, :, 2].all()) and not bool(out.mask[0].any())
    assert pos.shape == (2, 256, 2, 3) and rel_err(pos, O.image_sine_pos(out.mask)) < 1e-6
    # the reference's requires_grad policy (backbone.py:78-86): only layer2-4 convolutions train
    trainable = {k for k, p in enc.named_parameters() if p.requires_grad}
    assert trainable and all(("layer2" in k or "layer3" in k or "layer4" in k) for k in trainable)
    # gradients flow through the folded weights
    out.tensors.float().pow(2).mean().backward()
    assert enc[0].body["layer4"][0].conv1.weight.grad is not None and enc[0].body["conv1"].weight.grad is None


def test_feature_resizer_and_text_encoder_wrapper():
    from transformers import RobertaConfig, RobertaModel
    from stcat_b200.text import TextEncoder

    torch.manual_seed(2)
    body = RobertaModel(RobertaConfig(vocab_size=100, hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                                      intermediate_size=128, max_position_embeddings=40, type_vocab_size=1, pad_token_id=1)).eval()

    class Tok(dict):
        def to(self, device):
            return self

    class FakeTokenizer:
        def batch_encode_plus(self, texts, padding="longest", return_tensors="pt"):
            L = max(len(t.split()) for t in texts) + 2
            ids = torch.ones(len(texts), L, dtype=torch.long)
            att = torch.zeros(len(texts), L, dtype=torch.long)
            for i, t in enumerate(texts):
                n = len(t.split()) + 2
                ids[i, :n] = torch.arange(3, 3 + n)
                att[i, :n] = 1
            return Tok(input_ids=ids, attention_mask=att)

    enc = TextEncoder(outdim=256, body=body, tokenizer=FakeTokenizer()).eval()
    assert set(k for k in enc.state_dict() if k.startswith("resizer")) == {"resizer.fc.weight", "resizer.fc.bias",
                                                                            "resizer.layer_norm.weight", "resizer.layer_norm.bias"}
    texts = ["a man walks to the door", "the dog"]
    (mask, mem, tok), cls = enc(texts)
    with torch.no_grad():  # the reference's order: resizer on the sequence-first memory, then on the pooled vector (bert.py:63-77)
        e = body(input_ids=tok["input_ids"], attention_mask=tok["attention_mask"])
        P = {"r." + k[len("resizer."):]: v for k, v in enc.state_dict().items() if k.startswith("resizer")}
        mem_ref = O.feature_resizer(P, "r", e.last_hidden_state.transpose(0, 1))
        cls_ref = O.feature_resizer(P, "r", e.pooler_output)
    assert mem.shape == (8, 2, 256) and cls.shape == (2, 256) and mask.dtype == torch.bool
    assert torch.equal(mask, tok["attention_mask"].ne(1))
    assert rel_err(mem, mem_ref) < 1e-5 and rel_err(cls, cls_ref) < 1e-5
    # hoisted tokenisation: forward also takes the token tensors
    (_, mem2, _), _ = enc(enc.tokenize(texts))
    assert torch.equal(mem, mem2)


def test_double_pass_equals_two_passes_and_reference_merge():
    """evaluate.double_pass (one ragged forward over [even, odd] x clips + device post-processing / interpolation) against the
    reference's procedure restated: two separate forwards per clip, PostProcess, python linear_interp, segment union."""
    from stcat_b200 import synthetic
    from stcat_b200.evaluate import double_pass
    from stcat_b200.param_spec import synthetic_params
    from stcat_b200.pipeline import PostProcess, STCATHotPath

    cfg = cfg_for({"max_video_len": 16})
    model = STCATHotPath(cfg).load_flat_params(synthetic_params(cfg, seed=0)).eval()
    durations = [7, 4]
    inp = synthetic.make_inputs(durations, 3, 3, 4, seed=5)
    videos = NestedTensor(inp["vis_features"], inp["vis_mask"], durations)
    texts = (inp["text_mask"], inp["text_memory"], None)
    targets = [{"item_id": 11, "ori_size": (240, 320), "frame_ids": [3, 5, 8, 9, 12, 16, 17], "qtype": "declar"},
               {"item_id": 12, "ori_size": (100, 200), "frame_ids": [0, 2, 4, 7]}]
    run = lambda v, p, t: model(v, p, t)
    bbox, temp = double_pass(run, videos, inp["vis_pos"], texts, targets)
    post = PostProcess()
    offs = [0, 7]
    for i, tg in enumerate(targets):
        preds, steds = [], []
        for start in (0, 1):
            sl = slice(offs[i] + start, offs[i] + durations[i], 2)
            v1 = NestedTensor(inp["vis_features"][sl], inp["vis_mask"][sl], [len(range(*sl.indices(100)))])
            t1 = (inp["text_mask"][i:i + 1], inp["text_memory"][:, i:i + 1], None)
            with torch.no_grad():
                out = model(v1, inp["vis_pos"][sl], t1)
            fids = tg["frame_ids"][start::2]
            sizes = torch.tensor([list(tg["ori_size"])] * len(fids), dtype=torch.float32)
            bx, st = post(out, sizes, [fids], v1.durations)
            preds.append({f: [bx[j].tolist()] for j, f in enumerate(fids)})
            steds.append(st[0])
        ref_boxes, ref_sted = O.merge_even_odd(preds[0], steds[0], preds[1], steds[1])
        assert sorted(bbox[tg["item_id"]]) == sorted(ref_boxes)
        assert max(ref_boxes) - min(ref_boxes) + 1 == len(ref_boxes)  # evaluate.py:37
        for f in ref_boxes:
            a, r = torch.tensor(bbox[tg["item_id"]][f][0]), torch.tensor(ref_boxes[f][0])
            assert float((a - r).abs().max()) < 1e-3 * max(1.0, float(r.abs().max())), (tg["item_id"], f)
        assert temp[tg["item_id"]]["sted"] == ref_sted
    assert temp[11]["qtype"] == "declar" and "qtype" not in temp[12]

"""-m gpu: the kernels / modules either side of the hot path on the B200 through the C ABI."""
import pytest
import torch

from oracle import stcat_oracle as O
from helpers import cfg_for, rel_err
from stcat_b200 import ops
from stcat_b200.nested import NestedTensor

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def cuda_backend():
    ops.set_backend(None)
    ops.set_precision("fp32")
    ops.clear_weight_cache()
    yield
    ops.set_precision("fp32")
    ops.clear_weight_cache()


def test_pos_sine_kernel():
    from stcat_b200.vision import PositionEmbeddingSine

    for (n, H, W) in [(3, 7, 9), (64, 14, 14), (5, 14, 23), (2, 20, 20)]:
        mask = torch.zeros(n, H, W, dtype=torch.bool)
        if n > 1:
            mask[1, H - 2:, :] = True
            mask[n - 1, :, W - 3:] = True
        pos = PositionEmbeddingSine(128)(NestedTensor(torch.zeros(n, 256, H, W, device="cuda"), mask.cuda(), [n]))
        assert rel_err(pos, O.image_sine_pos(mask)) < 2e-5  # sinf / cosf / powf vs torch's CPU libm
    m0 = torch.zeros(64, 14, 14, dtype=torch.bool, device="cuda")
    pos = PositionEmbeddingSine(128)(NestedTensor(torch.zeros(64, 256, 14, 14, device="cuda"), m0, [64]))
    assert pos.shape == (64, 256, 14, 14) and rel_err(pos, O.image_sine_pos(m0.cpu())) < 2e-5


def test_box_interp_kernel():
    be = ops.get_backend()
    gen = torch.Generator().manual_seed(0)
    for ids in ([3, 5, 8, 9, 12, 16, 17], [0, 1, 2], [4, 40], [7]):
        boxes = torch.rand(len(ids), 4, generator=gen) * 300
        ref = O.linear_interp({f: [boxes[j].tolist()] for j, f in enumerate(ids)})
        first, last = ids[0], ids[-1]
        out = torch.empty(last - first + 1 + 2, 4, device="cuda")
        be.box_interp(torch.tensor(ids, device="cuda"), boxes.cuda(), out, first - 1)  # one frame outside on either side
        out = out.cpu()
        assert float(out[0, 0]) == -1.0 and float(out[-1, 0]) == -1.0
        for f, b in ref.items():
            assert float((out[f - first + 1] - torch.tensor(b[0])).abs().max()) < 1e-3


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 1e-2)])
def test_input_proj_on_device(precision, tol):
    from stcat_b200.vision import InputProj

    ops.set_precision(precision)
    torch.manual_seed(0)
    proj = InputProj(2048, 256).cuda()
    conv = torch.nn.Conv2d(2048, 256, 1).double()
    conv.load_state_dict({k: v.double().cpu() for k, v in proj.state_dict().items()})
    x = torch.randn(12, 2048, 7, 9).contiguous(memory_format=torch.channels_last)
    xd = (x.cuda().to(torch.bfloat16) if precision == "bf16" else x.cuda()).requires_grad_(True)
    xr = (x.to(torch.bfloat16).double() if precision == "bf16" else x.double()).requires_grad_(True)
    if precision == "bf16":
        with torch.no_grad():
            conv.weight.copy_(conv.weight.to(torch.bfloat16).double())
    y, yr = proj(xd), conv(xr)
    assert y.dtype == torch.float32 and rel_err(y, yr) < tol
    g = torch.randn(yr.shape)
    y.backward(g.cuda())
    yr.backward(g.double() if precision == "fp32" else g.to(torch.bfloat16).double())
    assert rel_err(proj.weight.grad, conv.weight.grad) < max(tol, 2e-5) * 3
    assert rel_err(proj.bias.grad, conv.bias.grad) < max(tol, 2e-5) * 3
    assert rel_err(xd.grad, xr.grad) < max(tol, 2e-5) * 3


def test_feature_resizer_on_device():
    from stcat_b200.text import FeatureResizer

    torch.manual_seed(1)
    rs = FeatureResizer(768, 256, dropout=0.1).cuda().eval()
    x = torch.randn(16, 2, 768)
    P = {"r." + k: v.cpu() for k, v in rs.state_dict().items()}
    assert rel_err(rs(x.cuda()), O.feature_resizer(P, "r", x)) < 1e-4


def test_double_pass_on_device_matches_cpu_emulation():
    from emu_backend import EmuBackend
    from stcat_b200 import synthetic
    from stcat_b200.evaluate import double_pass
    from stcat_b200.param_spec import synthetic_params
    from stcat_b200.pipeline import STCATHotPath

    cfg = cfg_for({"max_video_len": 16})
    P = synthetic_params(cfg, seed=0)
    durations = [7, 4]
    inp = synthetic.make_inputs(durations, 3, 3, 4, seed=5)
    targets = [{"item_id": 11, "ori_size": (240, 320), "frame_ids": [3, 5, 8, 9, 12, 16, 17], "qtype": "declar"},
               {"item_id": 12, "ori_size": (100, 200), "frame_ids": [0, 2, 4, 7]}]

    def run(device):
        model = STCATHotPath(cfg).load_flat_params(P).to(device).eval()
        mv = lambda t: t.to(device)
        videos = NestedTensor(mv(inp["vis_features"]), mv(inp["vis_mask"]), durations)
        return double_pass(lambda v, p, t: model(v, p, t), videos, mv(inp["vis_pos"]), (mv(inp["text_mask"]), mv(inp["text_memory"]), None), targets)

    bbox_g, temp_g = run("cuda")
    ops.set_backend(EmuBackend())
    try:
        bbox_c, temp_c = run("cpu")
    finally:
        ops.set_backend(None)
    assert temp_g == temp_c
    for vid in bbox_c:
        assert sorted(bbox_g[vid]) == sorted(bbox_c[vid])
        for f in bbox_c[vid]:
            a, r = torch.tensor(bbox_g[vid][f][0]), torch.tensor(bbox_c[vid][f][0])
            assert float((a - r).abs().max()) < 1e-2 * max(1.0, float(r.abs().max()))


def test_full_model_outer_seam_runs_in_bf16():
    """STCATNet (frames + captions in, prediction dict out): channels-last bf16 trunk with folded FrozenBN -> input_proj GEMM ->
    hot path, forward + backward; the trunk's bf16 features stay close to the fp32 trunk's."""
    from transformers import RobertaConfig, RobertaModel
    from stcat_b200.pipeline import STCATNet

    class Tok(dict):
        def to(self, device):
            return Tok({k: v.to(device) for k, v in self.items()})

    class FakeTokenizer:
        def batch_encode_plus(self, texts, padding="longest", return_tensors="pt"):
            L = 8
            return Tok(input_ids=torch.arange(3, 3 + L).repeat(len(texts), 1), attention_mask=torch.ones(len(texts), L, dtype=torch.long))

    torch.manual_seed(3)
    cfg = cfg_for({"max_video_len": 16})
    body = RobertaModel(RobertaConfig(vocab_size=100, hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                                      intermediate_size=128, max_position_embeddings=40, type_vocab_size=1, pad_token_id=1))
    ops.set_precision("bf16")
    net = STCATNet(cfg, text_body=body, tokenizer=FakeTokenizer()).cuda().eval()
    with torch.no_grad():  # a randomly initialised 101-layer trunk with identity BN emits features of magnitude 1e5; damp the
        for k, v in net.vis_encoder.state_dict().items():  # residual branches (as trained / zero-init-residual networks do)
            if k.endswith("bn3.weight"):
                v.fill_(0.1)
    frames = torch.randn(6, 3, 128, 160, device="cuda")
    videos = NestedTensor(frames, torch.zeros(6, 128, 160, dtype=torch.bool, device="cuda"), [6])
    f32, _ = net.vis_encoder(videos)
    net.vis_encoder.set_compute_dtype(torch.bfloat16)
    f16, pos = net.vis_encoder(videos)
    assert f16.tensors.dtype == torch.bfloat16 and f16.tensors.is_contiguous(memory_format=torch.channels_last)
    assert rel_err(f16.tensors.float(), f32.tensors) < 5e-2
    out = net(videos, ["a man walks to the door"])
    assert out["pred_boxes"].shape == (6, 4) and out["pred_sted"].shape == (1, 6, 2) and len(out["aux_outputs"]) == 5
    (out["pred_boxes"].sum() + out["pred_sted"].sum()).backward()
    assert net.input_proj.weight.grad is not None and torch.isfinite(net.input_proj.weight.grad).all()
    g = net.vis_encoder[0].body["layer4"][0].conv1.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0

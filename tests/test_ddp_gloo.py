"""CPU (-m "not gpu"), world_size 2 over gloo: the N>1 path of the hot path (DESIGN.md "Multi-GPU").

Each rank runs fwd + loss + bwd of its own clip with the wgrad kernels accumulating into the flat gradient buffer
(stcat_b200/dp.py), then ONE all-reduce of that buffer.  The result must equal the sum of the two clips' gradients
computed in a single process, and parameters the forward never uses (fusion.*, ca_qtime_proj.*) must stay zero
without any unused-parameter search (the reference needs find_unused_parameters=True, train_net.py:34).
The kernels are emulated on CPU (tests/emu_backend.py); the kernels themselves are covered by the -m gpu tests.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _clip_grads(rank_seed, fused_flat, overlapped=False, precision="fp32"):
    from helpers import cfg_for
    from emu_backend import EmuBackend
    from stcat_b200 import ops, synthetic
    from stcat_b200.dp import FlatGrads, GradSync, hot_path_groups
    from stcat_b200.loss import STGLossPlan
    from stcat_b200.nested import NestedTensor
    from stcat_b200.param_spec import synthetic_params
    from stcat_b200.pipeline import STCATHotPath

    ops.set_backend(EmuBackend())
    ops.set_precision(precision)
    ops.clear_weight_cache()
    cfg = cfg_for({"max_video_len": 16})
    model = STCATHotPath(cfg).load_flat_params(synthetic_params(cfg, seed=0)).eval()  # same weights on every rank
    T = 5
    inp = synthetic.make_inputs([T], 3, 3, 4, seed=rank_seed)  # rank-seeded clip (bench.py: seed = 42 + rank)
    tg = synthetic.make_targets([T], seed=rank_seed)
    plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], [T], "cpu")
    grads = FlatGrads(model, hot_path_groups(model)) if fused_flat else None
    sync = GradSync(grads) if overlapped else None
    if sync is not None:
        sync.prepare(model)
    ops.set_grad_fusion(fused_flat)
    try:
        vis = inp["vis_features"].clone().requires_grad_(True)
        out = model(NestedTensor(vis, inp["vis_mask"], [T]), inp["vis_pos"], (inp["text_mask"], inp["text_memory"], None))
        total, _ = plan(out)
        if sync is not None:
            # bucketed all-reduce from autograd hooks DURING backward (bench.py's mode): a range reduced before its
            # gradients were complete would show up as a mismatch against the single-process sums below
            assert sync.active
            sync.begin_step()
            sync.attach(out)
        total.backward()
        if sync is not None:
            order = list(sync.done)
            sync.finish()
            assert "decoder" in order and "enc0" in order, order  # the hooks fired (not just finish())
    finally:
        ops.set_grad_fusion(False)
        ops.set_precision("fp32")
    return model, grads, float(total.detach())


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (a) one all-reduce of the whole flat buffer after backward
        model_a, grads_a, loss_a = _clip_grads(42 + rank, fused_flat=True)
        local = grads_a.buf.clone()
        grads_a.all_reduce()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.equal(grads_a.buf, sum(gathered))
        # (b) bucketed all-reduce overlapped with backward: same numbers
        model, grads, loss = _clip_grads(42 + rank, fused_flat=True, overlapped=True)
        assert loss == loss_a
        # GradSync averages over the ranks (what the DistributedDataParallel wrapper it replaces does); (a) summed
        assert torch.allclose(grads.buf * world, grads_a.buf, rtol=1e-6, atol=1e-7)
        assert set(grads.ranges) >= {"decoder", "enc0", "enc5", "rest"}
        # (c) bf16 operand mode: fused layout glue (the decoder reads the encoder stream itself) + gradient sink: the decoder
        # range is still reduced from a hook during backward (asserted inside), and the result is the mean of the local sums
        _, grads_c0, loss_c0 = _clip_grads(42 + rank, fused_flat=True, precision="bf16")
        local_c = grads_c0.buf.clone()
        gathered_c = [torch.zeros_like(local_c) for _ in range(world)]
        dist.all_gather(gathered_c, local_c)
        _, grads_c, loss_c = _clip_grads(42 + rank, fused_flat=True, overlapped=True, precision="bf16")
        assert loss_c == loss_c0
        assert torch.allclose(grads_c.buf * world, sum(gathered_c), rtol=1e-5, atol=1e-6)
        # every .grad is still a view of the reduced buffer
        for p in grads.params:
            assert p.grad.untyped_storage().data_ptr() == grads.buf.untyped_storage().data_ptr()
        # max-over-ranks timing idiom of bench.py works over gloo too
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        named = {k: p.grad.clone() for k, p in model.named_parameters()}
        torch.save({"loss": loss, "grads": named}, os.path.join(outdir, f"rank{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_gradient_allreduce_world2(tmp_path):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0
    res = {}
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        res[r] = (d["loss"], d["grads"])
    # both ranks hold the same reduced gradients
    for k, g in res[0][1].items():
        assert torch.equal(g, res[1][1][k]), k
    # ... equal to the single-process MEAN of the two clips' ordinary autograd gradients (DDP semantics)
    sys.path.insert(0, HERE)
    singles = [_clip_grads(42 + r, fused_flat=False) for r in range(world)]
    from stcat_b200 import ops

    ops.set_backend(None)
    assert res[0][0] != res[1][0]  # different clips per rank
    for r in range(world):
        assert abs(singles[r][2] - res[r][0]) < 1e-6 * abs(singles[r][2])
    unused = 0
    for k, g in res[0][1].items():
        parts = [dict(m.named_parameters())[k].grad for m, _, _ in singles]
        if all(p is None for p in parts):
            assert float(g.abs().max()) == 0.0, k  # unused parameter: zero slice, no search needed
            unused += 1
            continue
        ref = sum(p for p in parts if p is not None) / world
        err = float((g - ref).abs().max())
        assert err <= 1e-5 * float(ref.abs().max()) + 1e-6, (k, err)  # abs floor: gradients that are zero in exact arithmetic (key-bias of a softmax) are fp32 noise
    assert unused >= 8  # fusion.{weight,bias} + 6 x ca_qtime_proj.{weight,bias} at least


def _ddp_worker(rank, world, port, outdir):
    """the reference's own wrapping (train_net.py:31-36): torch DistributedDataParallel(find_unused_parameters=True) around
    modules whose arithmetic is stcat_b200's autograd Functions; gradients must come out averaged over the ranks"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import cfg_for
        from emu_backend import EmuBackend
        from stcat_b200 import ops, synthetic
        from stcat_b200.loss import STGLossPlan
        from stcat_b200.nested import NestedTensor
        from stcat_b200.param_spec import synthetic_params
        from stcat_b200.pipeline import STCATHotPath

        ops.set_backend(EmuBackend())
        ops.set_precision("fp32")
        cfg = cfg_for({"max_video_len": 16})
        model = STCATHotPath(cfg).load_flat_params(synthetic_params(cfg, seed=0)).eval()
        ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
        T = 5
        inp = synthetic.make_inputs([T], 3, 3, 4, seed=42 + rank)
        tg = synthetic.make_targets([T], seed=42 + rank)
        plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], [T], "cpu")
        out = ddp(NestedTensor(inp["vis_features"], inp["vis_mask"], [T]), inp["vis_pos"], (inp["text_mask"], inp["text_memory"], None))
        total, _ = plan(out)
        total.backward()
        named = {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()}
        torch.save({"loss": float(total.detach()), "grads": named}, os.path.join(outdir, f"ddp_rank{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_torch_ddp_wrapping_world2(tmp_path):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), f"ddp_rank{r}.pt")) for r in range(world)]
    sys.path.insert(0, HERE)
    singles = [_clip_grads(42 + r, fused_flat=False) for r in range(world)]
    from stcat_b200 import ops

    ops.set_backend(None)
    checked = 0
    for k, g0 in res[0]["grads"].items():
        g1 = res[1]["grads"][k]
        parts = [dict(m.named_parameters())[k].grad for m, _, _ in singles]
        if all(p is None for p in parts):  # never used (fusion.*, ca_qtime_proj.*): DDP leaves None or zeros
            assert g0 is None or float(g0.abs().max()) == 0.0, k
            continue
        assert torch.equal(g0, g1), k  # both ranks hold the same averaged gradient
        mean = sum(p for p in parts if p is not None) / world
        # thread count / summation order differ between the workers and this process: fp32 noise only
        err = float((g0 - mean).abs().max())
        assert err <= 1e-4 * float(mean.abs().max()) + 1e-6, (k, err)
        checked += 1
    assert checked > 300

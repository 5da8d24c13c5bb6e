"""Diagnosis: which way of capturing an NCCL all-reduce into a CUDA graph works here (run under torchrun, 2 ranks)."""
import os, sys, traceback
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
buf = torch.ones(1 << 20, device=dev)
dist.all_reduce(buf); torch.cuda.synchronize()

def case(name, fn, mode):
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        torch.cuda.synchronize()
        buf.fill_(1.0); g.replay(); torch.cuda.synchronize()
        if rank == 0: print(f"{name:40s} mode={mode:12s} OK   buf[0]={float(buf[0]):.1f}", flush=True)
    except Exception as e:
        if rank == 0: print(f"{name:40s} mode={mode:12s} FAIL {type(e).__name__}: {str(e).splitlines()[0][:150]}", flush=True)
        torch.cuda.synchronize()

side = torch.cuda.Stream()
def plain(): dist.all_reduce(buf)
def on_side():
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side): dist.all_reduce(buf)
    torch.cuda.current_stream().wait_stream(side)
w = torch.ones(4, device=dev, requires_grad=True)
def from_hook():
    x = w * 2
    x.register_hook(lambda g: (on_side(), None)[1])
    (x.sum() * buf[0]).backward()
for mode in ("global", "thread_local", "relaxed"):
    case("plain current stream", plain, mode)
    case("side stream", on_side, mode)
    case("side stream from autograd hook", from_hook, mode)
dist.destroy_process_group()

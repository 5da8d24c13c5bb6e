#!/usr/bin/env python
"""Turns the raw captures under gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed):
  launches.csv (ncu --metrics gpu__time_duration.sum)  ->  profiles/<tag>_ncu_launches.md
  prof_*.ncu-rep (ncu --set full)                      ->  profiles/<tag>_ncu_<name>.md  (+ dominant_kernel_traffic.json)
Usage: python scripts/summarize_profiles.py <tag>"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launches():
    path = os.path.join(G, "launches.csv")
    if not os.path.exists(path):
        return
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        rows.append((r[ki], us))
    # the command ran 3 warm-up steps + 1 timed step eagerly: keep the last quarter (one step)
    n = len(rows) // 4
    step = rows[-n:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, us in step:
        k = re.sub(r"\(.*", "", k)[:110]
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{tag}_ncu_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of one eager step (bf16, T=64 res=448 L=16)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline`\n"
                f"(last quarter of {len(rows)} profiled launches = one step: {n} launches, {tot / 1e3:.2f} ms of serialised, cold-cache kernel time; compare SHARES)\n\n"
                "| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write(f"| `{k}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |\n")
    print("wrote launches summary:", n, "launches/step", round(tot / 1e3, 2), "ms")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def full(rep, name, note):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return None
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    vals = {}
    for i, h in enumerate(hdr):
        if h in WANT:
            vals[h] = (r[i], units[i])
    kname = r[hdr.index("Kernel Name")]
    with open(os.path.join(P, f"{tag}_ncu_{name}.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full, {name}\n\n{note}\n\nkernel: `{kname[:160]}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for h in WANT:
            if h in vals:
                f.write(f"| {h} | {vals[h][0]} | {vals[h][1]} |\n")
    print("wrote", name, {k: v[0] for k, v in vals.items() if k.startswith(("gpu__time", "dram__bytes", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"))})
    return vals


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


launches()
v = full("prof_ffn1.ncu-rep", "ffn1_gemm", "`ncu --set full --clock-control none -k regex:gemm_tc -s 3 -c 1 python scripts/bench_gemm.py --only ffn1_fwd --iters 2 --no-graph`: "
         "FFN linear1 forward (M=13632, N=2048, K=256, bf16 out, bias+ReLU).  Algorithmic bytes 63.9 MB (A 7.0 + W 1.0 + C 55.8); the "
         "output stays in the 126 MB L2 at kernel end, so DRAM traffic is below that.")
if v and "dram__bytes_read.sum" in v:
    t = to_bytes(*v["dram__bytes_read.sum"]) + to_bytes(*v["dram__bytes_write.sum"])
    json.dump({"kernel": "gemm_tc_kernel FFN linear1 fwd", "traffic_bytes_per_launch": t, "source": f"profiles/{tag}_ncu_ffn1_gemm.md (dram__bytes_read.sum + dram__bytes_write.sum)"},
              open(os.path.join(P, "dominant_kernel_traffic.json"), "w"))
full("prof_attn_bwd.ncu-rep", "attn_bwd", "`ncu --set full --clock-control none -k regex:attn_tc_bwd -c 1 python bench.py --steps 1 --warmup 3 --no-graph`: spatial encoder attention "
     "core backward (S, dP, dV, dK, dQ on tcgen05), 64 frames x 8 heads x (213 x 213 x 32), bf16.")
full("prof_attn_fwd.ncu-rep", "attn_fwd", "`ncu --set full --clock-control none -k regex:attn_tc_fwd -c 1 python bench.py --steps 1 --warmup 3 --no-graph`: spatial encoder attention "
     "core forward, 64 frames x 8 heads x (213 x 213 x 32), bf16.")

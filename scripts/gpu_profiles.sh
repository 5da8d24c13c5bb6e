#!/bin/bash
# Evidence for profiles/: smoke, default bench line (with cpu_baseline), reference arm, ncu launch list of one eager step,
# ncu --set full of the dominant kernel (FFN linear1 GEMM) and of the encoder attention kernels.  Every command under
# its own timeout.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
STCAT_NO_PDL=1 STCAT_TRACE=gpurun_out/trace_nopdl.json timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile gpurun_out/profile_nopdl.md > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "nopdl rc=$?"
gzip -f gpurun_out/trace_nopdl.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/prof_ffn1 \
    python scripts/bench_gemm.py --only ffn1_fwd --iters 2 --no-graph > gpurun_out/ncu_ffn1.log 2>&1; echo "ncu ffn1 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -c 1 -f -o gpurun_out/prof_attn_fwd \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn fwd rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -c 1 -f -o gpurun_out/prof_attn_bwd \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_attn_bwd.log 2>&1; echo "ncu attn bwd rc=$?"
timeout 100 python scripts/bench_gemm.py > gpurun_out/gemm_table.txt 2>&1; echo "gemm table rc=$?"

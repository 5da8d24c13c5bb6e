#!/bin/bash
# First GPU call of the next round: validate what was staged without GPU time at the end of round 1.
#   gpurun --timeout 300 -- 'bash scripts/validate_staged.sh'
# 1. the staged forward (STCAT_ATTN_FWD_V2=1: integer bf16 pack + O in its own TMEM columns) against the same parity tests
#    as the default kernel, and its timeline / graph-timed duration next to the default's;
# 2. the backward timeline and the GEMM timeline (never measured yet).
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== staged kernels that have never run on the GPU (fused optimizer step)"
timeout 120 python -m pytest tests -x -q -m gpu_staged 2>&1 | tail -3
echo "== default forward: parity + timeline"
timeout 120 python -m pytest tests/test_gpu_attention_tc.py -x -q 2>&1 | tail -2
timeout 60 python scripts/attn_timeline.py 64 213 fwd | tee gpurun_out/attn_fwd_timeline_v1.txt | tail -9
echo "== staged forward V2: parity + timeline"
STCAT_ATTN_FWD_V2=1 timeout 120 python -m pytest tests/test_gpu_attention_tc.py -x -q 2>&1 | tail -2
STCAT_ATTN_FWD_V2=1 timeout 60 python scripts/attn_timeline.py 64 213 fwd | tee gpurun_out/attn_fwd_timeline_v2.txt | tail -9
echo "== staged forward V2 + token-staggered softmax groups: parity + timeline"
STCAT_ATTN_FWD_V2=2 timeout 120 python -m pytest tests/test_gpu_attention_tc.py -x -q 2>&1 | tail -2
STCAT_ATTN_FWD_V2=2 timeout 60 python scripts/attn_timeline.py 64 213 fwd | tee gpurun_out/attn_fwd_timeline_v2s.txt | tail -9
echo "== staged backward (integer bf16 pack): parity + timing"
STCAT_ATTN_BWD_V2=1 timeout 120 python -m pytest tests/test_gpu_attention_tc.py -x -q -s -k "bf16_fwd_bwd or timing or tcgen05" 2>&1 | grep -E "passed|failed|us/launch" | tail -6
echo "== backward timeline"
timeout 60 python scripts/attn_timeline.py 64 213 bwd | tee gpurun_out/attn_bwd_timeline.txt | tail -20
echo "== GEMM timeline (FFN linear1 forward; FFN linear2 forward)"
timeout 60 python scripts/gemm_timeline.py 13632 2048 256 | tee gpurun_out/gemm_timeline_ffn1.txt | tail -20
timeout 60 python scripts/gemm_timeline.py 13632 256 2048 | tee gpurun_out/gemm_timeline_ffn2.txt | tail -8
echo "== staged GEMM epilogue with two staging tiles: parity + timeline + shape table"
STCAT_GEMM_EPI2=1 timeout 200 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -2
STCAT_GEMM_EPI2=1 timeout 60 python scripts/gemm_timeline.py 13632 2048 256 | tee gpurun_out/gemm_timeline_ffn1_epi2.txt | tail -10
timeout 120 python scripts/bench_gemm.py 2>/dev/null | head -12
STCAT_GEMM_EPI2=1 timeout 120 python scripts/bench_gemm.py 2>/dev/null | head -12
echo "== graph-timed attention core, default vs V2"
for v in "" 1 2; do
  if [ -n "$v" ]; then export STCAT_ATTN_FWD_V2=$v; fi
  timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read()); e = d['encoder_attention']
print('V2 level $v' if '$v' else 'V1', 'step ms', round(d['ms_per_step'], 3), 'core us', round(e['us_core'], 2), 'block us', round(e['us_block'], 2))"
done

#!/bin/bash
# compute-sanitizer over the hand-written kernels (the reference has no race detection; the tcgen05 / TMA kernels
# here hand-roll their mbarrier protocols, so they are checked on hardware):
#   memcheck  : out-of-bounds / misaligned global and shared accesses, including the TMA boxes at the M / N / K tails
#   racecheck : shared-memory hazards between the warp roles (staging tiles, bias slices, probability tiles)
#   synccheck : invalid barrier usage (bar.sync / mbarrier arrive counts)
# Small shapes only (the sanitizer serialises the kernels and runs 10-100x slower); each tool gets its own log under
# gpurun_out/ and the script prints one summary line per tool.  Usage: bash scripts/sanitize.sh [memcheck racecheck synccheck]
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
TOOLS=${@:-memcheck racecheck synccheck}
GEMM='test_fwd_bf16 and (128-256-64 or 500-96-72 or 64-256-2048 or 300-104-1024) or test_bwd_bf16 and (500-96-72 or 64-2048-256) or test_linear_group and 64-bf16 or test_bwd_data_fused_relu_mask_and_bias_grad and 300-64-192 or test_weight_resident_variant and 40000-264-72'
ATTN='test_attention_bf16_fwd_bwd and (2-8-213 or 5-8-66 or 3-8-257 or 1-8-129) or test_single_query_attention and (9-8-65 or 64-8-212-True) or test_tcgen05_attention_with_dropout and 2-8-213'
KERN='layernorm or anchor or sted or dropout or fused_loss'
for tool in $TOOLS; do
  log=gpurun_out/sanitize_$tool.log
  : > $log
  for sel in "tests/test_gpu_gemm_tc.py|$GEMM" "tests/test_gpu_attention_tc.py|$ATTN" "tests/test_gpu_kernels.py|$KERN"; do
    file=${sel%%|*}; k=${sel#*|}
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
      python -m pytest $file -q -x -m gpu -k "$k" -p no:cacheprovider >> $log 2>&1
    echo "[$tool] $file exit=$?" | tee -a $log
  done
  echo "[$tool] ERROR SUMMARY lines:"; grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $log | sort | uniq -c
done

#!/bin/bash
# weight-resident GEMM variant (BRES) + warp epilogue (WEPI): parity, phase timeline, table of the step's GEMM shapes
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x > gpurun_out/r2_w_pytest.log 2>&1; echo "pytest (default) rc=$?"; grep -v Warning gpurun_out/r2_w_pytest.log | tail -12
STCAT_GEMM_WEPI=1 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x > gpurun_out/r2_w_pytest_wepi.log 2>&1; echo "pytest (WEPI=1) rc=$?"; grep -v Warning gpurun_out/r2_w_pytest_wepi.log | tail -12
for cfg in "BRES=1" "BRES=0 WEPI=1" "BRES=0 WEPI=0"; do
  tag=$(echo $cfg | tr -d ' =')
  envs=$(for kv in $cfg; do echo -n "STCAT_GEMM_$kv "; done)
  echo "=== $envs timeline"
  env $envs timeout 120 python scripts/gemm_timeline.py 2>&1 | tail -22 | tee gpurun_out/r2_w_timeline_$tag.txt
  echo "=== $envs table"
  env $envs timeout 300 python scripts/bench_gemm.py --iters 30 2>&1 | tee gpurun_out/r2_w_gemm_table_$tag.txt | head -32
done

#!/bin/bash
# compute-sanitizer memcheck + racecheck over the tests of the layout-glue kernels (csrc/assembly.cu): transposes with row /
# column tails, frame-ordered reductions through shared memory, the box head
mkdir -p gpurun_out
export PYTHONPATH=.
K='token_assembly or mem_operands or template_generator or box_head'
for tool in memcheck racecheck; do
  log=gpurun_out/r2_ay_sanitize_$tool.log
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "$K" -p no:cacheprovider > $log 2>&1
  echo "[$tool] exit=$?"; grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $log | sort | uniq -c
done

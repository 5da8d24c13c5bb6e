#!/bin/bash
# 2 GPUs: SM cap for the persistent kernels while the gradient all-reduce is in flight
mkdir -p gpurun_out
export PYTHONPATH=.
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], ": ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as ex:
    print(sys.argv[2], "failed", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_aa_bench_1gpu.json 2> gpurun_out/r2_aa_bench_1gpu.err; show gpurun_out/r2_aa_bench_1gpu.json "1 GPU"
for cfg in "32 0" "16 16" "8 8" "16 0"; do
  set -- $cfg
  NCCL_MAX_CTAS=$1 STCAT_NCCL_CTAS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_aa_bench_2gpu_$1_$2.json 2> gpurun_out/r2_aa_bench_2gpu_$1_$2.err
  show gpurun_out/r2_aa_bench_2gpu_$1_$2.json "2 GPUs NCCL_MAX_CTAS=$1 cap=$2"
done

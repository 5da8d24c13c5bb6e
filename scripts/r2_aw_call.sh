#!/bin/bash
# 2-GPU check of the fused-glue / gradient-sink step with the NCCL gradient exchange; NCCL CTA budget A/B
mkdir -p gpurun_out
export PYTHONPATH=.
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], ": ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"])
except Exception as ex:
    print(sys.argv[2], "failed", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_aw_bench_1gpu.json 2> gpurun_out/r2_aw_bench_1gpu.err; show gpurun_out/r2_aw_bench_1gpu.json "N=1"
for ctas in 24 12 8; do
  NCCL_MAX_CTAS=$ctas timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((ctas % 10)) bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_aw_bench_2gpu_ctas$ctas.json 2> gpurun_out/r2_aw_bench_2gpu_ctas$ctas.err
  show gpurun_out/r2_aw_bench_2gpu_ctas$ctas.json "N=2 NCCL_MAX_CTAS=$ctas"
done

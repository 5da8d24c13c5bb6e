#!/bin/bash
# final scaling lines on one 8-GPU box: N = 1, 2, 4, 8 back to back (the driver's contract command)
mkdir -p gpurun_out
export PYTHONPATH=.
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], ": ms", round(d["ms_per_step"], 3), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"])
except Exception as ex:
    print(sys.argv[2], "failed", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final_bench_scale_1gpu.json 2> gpurun_out/r2_final_bench_scale_1gpu.err; show gpurun_out/r2_final_bench_scale_1gpu.json "N=1"
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final_bench_scale_${n}gpu.json 2> gpurun_out/r2_final_bench_scale_${n}gpu.err
  show gpurun_out/r2_final_bench_scale_${n}gpu.json "N=$n"
done

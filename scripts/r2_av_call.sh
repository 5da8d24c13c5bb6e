#!/bin/bash
# un-profiled phase durations of the captured step (external event nodes) for the A/B switches, one box
mkdir -p gpurun_out
export PYTHONPATH=.
run() {
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --phases > gpurun_out/r2_av_bench_$tag.json 2> gpurun_out/r2_av_bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_av_bench_$tag.json"))
    print("$tag step: ms", round(d["ms_per_step"], 3), "launches/step", d["gpu_launches"] / d["steps"])
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2_av_bench_$tag.err").read()[-1500:])
PY
grep phases gpurun_out/r2_av_bench_$tag.err
}
run default
run head0 STCAT_FUSED_HEAD=0
run group0 STCAT_GROUP_MEMSIDE=0
run head0_group0 STCAT_FUSED_HEAD=0 STCAT_GROUP_MEMSIDE=0
run single STCAT_SINGLE_STREAM=1

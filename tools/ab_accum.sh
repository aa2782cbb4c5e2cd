#!/bin/bash
# A/B timing of the accumulator-pass slicing (KB200_ACCUM_SLICE = poses whose accumulators are alive together) through bench.py's
# per-stage CUDA-event timers.
TAG=${1:-r02}
mkdir -p gpurun_out
for v in 32 4 8; do
  KB200_ACCUM_SLICE=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-pipeline > gpurun_out/ab_${TAG}_slice$v.json 2> gpurun_out/ab_${TAG}_slice$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${TAG}_slice$v.json"))
print("slice $v", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "accum_ms", round(d["stage_ms_per_launch"]["splat_accum"],4), "frac", round(d["roofline"]["frac"],3), {k:round(x,3) for k,x in d["stage_ms_per_launch"].items()})
PY
done

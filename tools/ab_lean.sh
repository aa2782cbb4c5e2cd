#!/bin/bash
# Parity tests of the conv kernels in both operand modes, then layer and network timings (KB_CONV_NO_LEAN=1 = generic epilogue).
mkdir -p gpurun_out
T="timeout 420"
TAG=${1:-lean}
$T python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
KB200_CONV_F16=1 $T python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
$T python tools/bench_conv.py --no-cudnn --nets > gpurun_out/${TAG}_tf32.jsonl 2> gpurun_out/${TAG}_tf32.err
$T python tools/bench_conv.py --no-cudnn --f16 > gpurun_out/${TAG}_f16_layers.jsonl 2> gpurun_out/${TAG}_f16_layers.err
KB200_CONV_F16=1 $T python tools/bench_conv.py --no-cudnn --nets > gpurun_out/${TAG}_f16.jsonl 2> gpurun_out/${TAG}_f16.err
python - <<PY
import json
for f in ("${TAG}_tf32", "${TAG}_f16_layers", "${TAG}_f16"):
    print(f)
    for l in open(f"gpurun_out/{f}.jsonl"):
        d = json.loads(l)
        print("  %-58s %8.1f us" % (d.get("case") or d.get("net"), d["ms"] * 1e3))
PY

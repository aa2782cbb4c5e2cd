#!/bin/bash
# Parity tests of the conv kernels in both operand modes, then layer and network timings for KB_CONV_TEAMS = 4 (default) and 2.
# Other switches: KB_CONV_ONE_ISSUER=1, KB_CONV_NO_LEAN=1 (generic epilogue), KB_CONV_NO_WIDE=1 (16-byte loads / stores).
mkdir -p gpurun_out
T="timeout -k 10 300"
TAG=${1:-lean}
$T python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
KB200_CONV_F16=1 $T python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
for teams in 4 2; do
  export KB_CONV_TEAMS=$teams
  $T python tools/bench_conv.py --no-cudnn --nets > gpurun_out/${TAG}_t${teams}_tf32.jsonl 2> gpurun_out/${TAG}_t${teams}_tf32.err
  $T python tools/bench_conv.py --no-cudnn --f16 > gpurun_out/${TAG}_t${teams}_f16_layers.jsonl 2> gpurun_out/${TAG}_t${teams}_f16_layers.err
  KB200_CONV_F16=1 $T python tools/bench_conv.py --no-cudnn --nets > gpurun_out/${TAG}_t${teams}_f16.jsonl 2> gpurun_out/${TAG}_t${teams}_f16.err
done

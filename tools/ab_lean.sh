#!/bin/bash
# Parity tests of the conv kernels in both operand modes, then layer and network timings with two MMA issuers (default) and one
# (KB_CONV_ONE_ISSUER=1).  Other switches: KB_CONV_NO_LEAN=1 generic epilogue, KB_CONV_TEAMS, KB_CONV_NO_WIDE=1 16-byte stores.
mkdir -p gpurun_out
T="timeout -k 10 300"
TAG=${1:-lean}
$T python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
KB200_CONV_F16=1 $T python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
for one in 0 1; do
  export KB_CONV_ONE_ISSUER=$one
  $T python tools/bench_conv.py --no-cudnn --nets > gpurun_out/${TAG}_i${one}_tf32.jsonl 2> gpurun_out/${TAG}_i${one}_tf32.err
  $T python tools/bench_conv.py --no-cudnn --f16 > gpurun_out/${TAG}_i${one}_f16_layers.jsonl 2> gpurun_out/${TAG}_i${one}_f16_layers.err
  KB200_CONV_F16=1 $T python tools/bench_conv.py --no-cudnn --nets > gpurun_out/${TAG}_i${one}_f16.jsonl 2> gpurun_out/${TAG}_i${one}_f16.err
done
grep -h '"net"' gpurun_out/${TAG}_i*_tf32.jsonl gpurun_out/${TAG}_i*_f16.jsonl | cut -c1-120

#!/bin/bash
# Which stage bounds the persistent conv kernel: layer timings with parts of the pipeline switched off (results are wrong).
mkdir -p gpurun_out
for dbg in 0 6; do
  for wide in 0 1; do
  for mode in "--f16" ""; do
    echo "== KB_CONV_DEBUG=$dbg KB_CONV_NO_WIDE=$wide $mode"
    for i in 0 4 1; do
      KB_CONV_NO_WIDE=$wide KB_CONV_DEBUG=$dbg timeout -k 5 120 python tools/bench_conv.py --no-cudnn --only $i $mode 2>/dev/null | cut -c1-90
    done
  done
  done
done

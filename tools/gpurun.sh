#!/bin/bash
# Rebuild libkb200.so + the oracle, then run a command on the GPU box:  tools/gpurun.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
make -s -C ken_burns_effect_b200/csrc > /dev/null
make -s -C oracle libkb_oracle.so > /dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"

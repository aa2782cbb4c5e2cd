#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + full captures of the hot kernels, conv and pipeline timings.
# Usage (on the GPU box, via gpurun):  bash tools/gpu_round.sh <tag>       -> everything lands in gpurun_out/
# Every command runs under its own timeout: a hung kernel must not hold the box until gpurun's limit.
TAG=${1:-r02}
T="timeout -k 10"
mkdir -p gpurun_out
$T 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multirank.py 2>&1 | tail -15 | grep -v "objectMatch\|SyntaxWarning" > gpurun_out/pytest_gpu_$TAG.txt; cat gpurun_out/pytest_gpu_$TAG.txt
KB200_CONV_F16=1 $T 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_reference_e2e.py -m gpu -q 2>&1 | tail -5 | grep -v "objectMatch\|SyntaxWarning" > gpurun_out/pytest_gpu_${TAG}_f16.txt; cat gpurun_out/pytest_gpu_${TAG}_f16.txt
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.txt 2>&1; tail -2 gpurun_out/smoke_$TAG.txt
$T 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; cat gpurun_out/bench_${TAG}_reference.json | cut -c1-400
$T 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
KB200_CONV_F16=1 $T 400 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_f16.json 2> gpurun_out/bench_${TAG}_f16.err; cat gpurun_out/bench_${TAG}_f16.json | cut -c1-300
$T 300 python bench.py --steps 10 --warmup 3 --dolly --no-full-pipeline > gpurun_out/bench_${TAG}_dolly.json 2> gpurun_out/bench_${TAG}_dolly.err; cat gpurun_out/bench_${TAG}_dolly.json | cut -c1-600
$T 300 python tools/bench_conv.py --nets > gpurun_out/bench_conv_$TAG.jsonl 2> gpurun_out/bench_conv_$TAG.err
$T 300 python tools/bench_conv.py --no-cudnn --f16 > gpurun_out/bench_conv_${TAG}_f16_layers.jsonl 2> gpurun_out/bench_conv_${TAG}_f16_layers.err
KB200_CONV_F16=1 $T 300 python tools/bench_conv.py --no-cudnn --nets > gpurun_out/bench_conv_${TAG}_f16.jsonl 2> gpurun_out/bench_conv_${TAG}_f16.err
$T 300 python tools/profile_pipeline.py > gpurun_out/pipeline_profile_$TAG.txt 2>&1
KB200_CONV_F16=1 $T 300 python tools/profile_pipeline.py > gpurun_out/pipeline_profile_${TAG}_f16.txt 2>&1
$T 300 python tools/kbe_breakdown.py > gpurun_out/kbe_breakdown_$TAG.json 2> gpurun_out/kbe_breakdown_$TAG.err
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-pipeline > gpurun_out/ncu_bench_$TAG.log 2>&1
$T 400 ncu --set full --clock-control none --import-source on -k regex:'kf_accum|kf_fill|kf_splat_min|kf_degrid|kf_resolve|kf_crop_resize' -s 12 -c 6 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-pipeline > gpurun_out/ncu_full_$TAG.log 2>&1
for i in 0 1 4; do
  $T 300 ncu --set full --clock-control none --import-source on -k regex:k_conv -s 3 -c 1 -o gpurun_out/prof_conv_${TAG}_$i -f \
      python tools/bench_conv.py --only $i --iters 2 --no-cudnn > gpurun_out/ncu_conv_${TAG}_$i.log 2>&1
  $T 300 ncu --set full --clock-control none --import-source on -k regex:k_conv -s 3 -c 1 -o gpurun_out/prof_conv_${TAG}_f16_$i -f \
      python tools/bench_conv.py --only $i --iters 2 --no-cudnn --f16 > gpurun_out/ncu_conv_${TAG}_f16_$i.log 2>&1
done
ls -la gpurun_out | tail -12

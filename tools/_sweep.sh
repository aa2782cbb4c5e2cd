for e in 0 1 2 7; do KB_ACCUM_EXP=$e python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-pipeline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('exp', $e, {k:round(v,4) for k,v in d['stage_ms_per_launch'].items() if k in ('splat_accum','resolve','memset_accum')})"; done

for pg in 4 2 1; do KB_ACCUM_PG=$pg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-pipeline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pg', $pg, round(d['value']), round(d['roofline']['frac'],3), {k:round(v,4) for k,v in d['stage_ms_per_launch'].items()})"; done

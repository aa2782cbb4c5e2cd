for b in 16 24 32; do python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline --no-full-pipeline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch', $b, round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3))"; done

python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -3
python tools/bench_conv.py --no-cudnn --nets | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if 'case' in d: print(d['case'], round(d['ms']*1000,1), round(d['frac_tf32_peak'],3))
    else: print(d['net'], round(d['ms'],3))"

for nb in 0 64 32 16; do python tools/bench_conv.py --no-cudnn --only 13 --n_block $nb | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n_block', $nb, d['case'], round(d['ms']*1000,1))"; done

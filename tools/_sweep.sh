for i in 0 4 2; do for dbg in 0 1 2 3; do KB_CONV_DEBUG=$dbg python tools/bench_conv.py --no-cudnn --only $i | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('debug', $dbg, d['case'], round(d['ms']*1000,1))"; done; done

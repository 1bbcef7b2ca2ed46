#!/bin/bash
# quick GPU check: bench value only (dev tool)
timeout 300 python bench.py --steps ${1:-5} --warmup 3 --no-cpu-baseline --generations 0 ${@:2} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print('value %.4g  ms %.3f  e2e %.4g  fp64frac %.3f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']))
"

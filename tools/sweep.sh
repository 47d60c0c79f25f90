#!/bin/bash
# usage: tools/sweep.sh VAR v1 v2 ...   -> quick bench (NT=301) per value of an engine env knob
var=$1; shift
for v in "$@"; do
  export $var=$v
  python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$var=$v', 'value %.0f'%d['value'], 'bwd %.1f us frac %.3f'%(1e3*r['avg_launch_ms'], r['frac']), 'fwd %.1f us frac %.3f'%(1e3*r['forward_step']['avg_launch_ms'], r['forward_step']['frac']))"
done

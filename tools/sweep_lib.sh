#!/bin/bash
# usage: tools/sweep_lib.sh name1 name2 ...  -> quick bench per kernel-variant library
for v in "$@"; do
  if [ "$v" = base ]; then unset RTM_LIB_PATH; else export RTM_LIB_PATH=$PWD/rtm_gpu_b200/build/variants/librtm_$v.so; fi
  python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v', 'value %.0f'%d['value'], 'bwd %.1f us frac %.3f'%(1e3*r['avg_launch_ms'], r['frac']), 'fwd %.1f us frac %.3f'%(1e3*r['forward_step']['avg_launch_ms'], r['forward_step']['frac']))"
done

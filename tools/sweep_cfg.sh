#!/bin/bash
# usage: tools/sweep_cfg.sh "c5" base v1 ...   -> perf_configs per kernel-variant library
cfg=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset RTM_LIB_PATH; else export RTM_LIB_PATH=$PWD/rtm_gpu_b200/build/variants/librtm_$v.so; fi
  echo "== $v"
  python tools/perf_configs.py $cfg 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'], '| %.0f Mcell/s | fwd %.1f us %.3f | bwd %.1f us %.3f' % (d['Mcell_updates_per_s'], d['fwd_us'], d['fwd_frac'], d['bwd_us'], d['bwd_frac']))"
done

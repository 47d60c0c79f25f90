#!/bin/bash
# usage: tools/kernel_time.sh REGEX name1 name2 ...  -> isolated durations (ncu, us) of kernels matching REGEX per variant library
rx=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset RTM_LIB_PATH; else export RTM_LIB_PATH=$PWD/rtm_gpu_b200/build/variants/librtm_$v.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$rx -s 4 -c 8 --csv python bench.py --nt 41 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
from collections import defaultdict
d=defaultdict(list)
for r in rows: d[r[4].split('(')[0][-28:]+' grid'+r[6]].append(float(r[-1].replace(',',''))/1e3)
for k,v in d.items(): print('$v', k, 'n=%d avg %.1f us min %.1f'%(len(v), sum(v)/len(v), min(v)))"
done

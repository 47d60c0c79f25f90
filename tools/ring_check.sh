#!/bin/bash
# tools/ring_check.sh -- one gpurun call after a kernel change: GPU suite, sanitizer on the small runs, quick timing.
out=gpurun_out/ring; mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
tools/sweep_lib.sh base "$@" > $out/sweep.txt 2>&1; cat $out/sweep.txt
timeout 300 python bench.py --nt 301 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --shots-per-step 32 2>/dev/null | tail -1 > $out/b32.json
python -c "
import json; d=json.loads(open('$out/b32.json').read()); r=d['roofline']; print('b32', round(d['value']), 'bwd %.1f fwd %.1f'%(1e3*r['avg_launch_ms'],1e3*r['forward_step']['avg_launch_ms']))"
for t in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $t python tools/sanitize_small.py > $out/sanitize_$t.log 2>&1
  echo "$t rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" $out/sanitize_$t.log | tail -3
done

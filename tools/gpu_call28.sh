#!/bin/bash
# round 2, GPU call 28: forward look-ahead distance (finer), segment length with it, other configurations
out=gpurun_out/c28; mkdir -p $out
q() { name=$1; shift; python bench.py "$@" --no-cpu-baseline --no-e2e --no-ref-cuda 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$name', d['config']['workload'], 'value %.0f'%d['value'], 'bwd %.1f us'%(1e3*r['avg_launch_ms']), 'fwd %.1f us'%(1e3*r['forward_step']['avg_launch_ms']), d['clocks'].get('sm_mhz'))"; }
A="--nt 301 --steps 3 --warmup 1"
for v in 222 296 333 370 444 592 888; do export RTM_LOOKAHEAD_F=$v; q "LOOKAHEAD_F=$v" $A; done 2>&1 | tee $out/sweep_lookahead_f.txt
export RTM_LOOKAHEAD_F=296
for v in 11 12 13 14; do export RTM_SEG_TILES=$v; q "LOOKAHEAD_F=296 SEG_TILES=$v" $A; done 2>&1 | tee $out/sweep_seg.txt
unset RTM_SEG_TILES
for v in 148 296 444; do export RTM_LOOKAHEAD_F=$v
  q "c4 LOOKAHEAD_F=$v" --config c4 --nt 400 --steps 2 --warmup 1
  q "c5 LOOKAHEAD_F=$v" --config c5 --steps 2 --warmup 1
  q "c5:8:taylor LOOKAHEAD_F=$v" --config c5:8:taylor --steps 2 --warmup 1
  q "c3 LOOKAHEAD_F=$v" --config c3 --nt 1000 --steps 2 --warmup 1
done 2>&1 | tee $out/sweep_other_configs.txt

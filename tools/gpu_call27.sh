#!/bin/bash
# round 2, GPU call 27: segment length of the streamed launches and forward look-ahead at 64 shots per launch (NT = 301, boost clocks)
out=gpurun_out/c27; mkdir -p $out
q() { python bench.py --nt 301 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-ref-cuda 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1', 'value %.0f'%d['value'], 'bwd %.1f us'%(1e3*r['avg_launch_ms']), 'fwd %.1f us'%(1e3*r['forward_step']['avg_launch_ms']), d['clocks'].get('sm_mhz'))"; }
for v in 8 4 5 6 7 10 12 16 8; do export RTM_SEG_TILES=$v; q "RTM_SEG_TILES=$v"; done 2>&1 | tee $out/sweep_seg_tiles.txt
unset RTM_SEG_TILES
for v in 148 0 74 296 444; do export RTM_LOOKAHEAD_F=$v; q "RTM_LOOKAHEAD_F=$v"; done 2>&1 | tee $out/sweep_lookahead_f.txt

#!/bin/bash
# tools/ncu_ring.sh -- ncu --set full with source counters of one forward launch and one ring+frame backward launch
out=gpurun_out/ncu; mkdir -p $out
B="python bench.py --nt 41 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:fwd_step_kernel -s 10 -c 1 -f -o $out/fwd $B > $out/fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_step_kernel -s 6 -c 1 -f -o $out/bwdring $B > $out/bwd.log 2>&1
ls -la $out

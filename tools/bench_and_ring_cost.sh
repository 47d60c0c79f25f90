#!/bin/bash
# tools/bench_and_ring_cost.sh -- default bench line of the final code + cost of the ring tiles at the default batch
out=gpurun_out/records2; mkdir -p $out
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tools/sweep_lib.sh base noring > $out/sweep_noring.txt 2>&1; cat $out/sweep_noring.txt
python -c "
import json; d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['clocks'], d['roofline']['dram'])"

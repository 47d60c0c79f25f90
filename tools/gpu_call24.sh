#!/bin/bash
# round 2, GPU call 24: the code as committed -- full GPU suite, smoke, default bench, reference arm
out=gpurun_out/c24; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -3 $out/smoke.log
( time timeout 1200 python bench.py > $out/bench_default.json 2> $out/bench_default.err )
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err )
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c24/bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print('default', round(d['value']), 'bwd', round(1e3*r['avg_launch_ms'],1), 'fwd', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'], 'e2e', round(d['e2e']['value']), 'parity', d.get('parity_checked'), 'dram frac', round(r['dram']['frac'],3), 'exec', round(r['executed']['frac'],3), 'ref_cuda', d.get('ref_cuda_baseline',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d.get('gpu_launches'))
d=json.loads(open('gpurun_out/c24/bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', d['value'])
PY

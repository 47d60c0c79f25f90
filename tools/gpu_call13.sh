#!/bin/bash
# round 2, GPU call 13: final code -- full GPU suite, smoke, default bench (64 shots per step), reference arm, final ncu captures
out=gpurun_out/c13; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -16 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -2 $out/smoke.log
( time timeout 1200 python bench.py > $out/bench_default.json 2> $out/bench_default.err )
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err )
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c13/bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print('default', round(d['value']), 'bwd', round(1e3*r['avg_launch_ms'],1), 'fwd', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'], 'e2e', round(d['e2e']['value']), 'parity', d.get('parity_checked'), 'dram frac', round(r['dram']['frac'],3), 'exec', round(r['executed']['frac'],3), 'ref_cuda', d.get('ref_cuda_baseline',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'))
d=json.loads(open('gpurun_out/c13/bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['config'].get('sample','')[:120])
PY
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream2 -s 2 -c 2 -o $out/prof_stream_bwd $P > $out/ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ring_kernel -s 4 -c 2 -o $out/prof_ring $P > $out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fwd_step -s 4 -c 1 -o $out/prof_fwd $P > $out/ncu_full3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:thin_frame -s 2 -c 1 -o $out/prof_thin $P > $out/ncu_full4.log 2>&1
for t in memcheck racecheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_stream.py > $out/$t.log 2>&1; tail -2 $out/$t.log; done

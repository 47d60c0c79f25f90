#!/bin/bash
# round 2, GPU call 20: final code -- full GPU suite, smoke, default bench, reference arm, launch list, ncu captures (summarised on the box), memcheck
out=gpurun_out/c20; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -16 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -3 $out/smoke.log
( time timeout 1200 python bench.py > $out/bench_default.json 2> $out/bench_default.err )
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err )
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c20/bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print('default', round(d['value']), 'bwd', round(1e3*r['avg_launch_ms'],1), 'fwd', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'], 'e2e', round(d['e2e']['value']), 'parity', d.get('parity_checked'), 'dram frac', round(r['dram']['frac'],3), 'exec', round(r['executed']['frac'],3), 'ref_cuda', d.get('ref_cuda_baseline',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d.get('gpu_launches'))
d=json.loads(open('gpurun_out/c20/bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', d['value'])
PY
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e --no-ref-cuda"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
python tools/ncu_summary.py launches $out/launches.csv > $out/launches.txt 2>&1; head -12 $out/launches.txt
cap() { name=$1; kern=$2; skip=$3; shift 3; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -o $out/prof_$name "$@" > $out/ncu_$name.log 2>&1
        python tools/ncu_summary.py full $out/prof_$name.ncu-rep > $out/${name}_full.txt 2>&1
        ncu -i $out/prof_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $out/${name}_source.csv.gz; rm -f $out/prof_$name.ncu-rep; }
cap stream_bwd stream2 2 $P
cap ring ring_kernel 4 $P
cap fwd fwd_step 4 $P
cap thin thin_frame 2 $P
for t in memcheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_stream.py > $out/$t.log 2>&1; tail -2 $out/$t.log; done
ls -la $out; du -sh gpurun_out

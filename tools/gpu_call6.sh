#!/bin/bash
# round 2, GPU call 6: ring_kernel (TMA boxes + per-model one-way coefficients): parity everywhere, A/B timing
out=gpurun_out/c6; mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x -k "not c5 and not 5000 and not fullsize" ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; ( env "$@" timeout 300 $B 2> $out/bench_$name.err | tail -1 > $out/bench_$name.json ); echo "$name rc=$?"; }
run ring0 RTM_RING2=0
run ring1_fwd0 RTM_RING2=1 RTM_RING2_FWD=0
run ring1_fwd1 RTM_RING2=1 RTM_RING2_FWD=1
run tile_ring1 RTM_STREAM2=0 RTM_RING2=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c6/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), 'exec frac', round(r['executed']['frac'],3))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ring_kernel -s 4 -c 2 -o $out/prof_ring $P > $out/ncu_full.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_stream.py > $out/memcheck.log 2>&1
tail -4 $out/memcheck.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $out/memcheck_small.log 2>&1
tail -4 $out/memcheck_small.log

#!/bin/bash
# round 2, GPU call 7: full GPU test suite + the default bench (full workload)
out=gpurun_out/c7; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=15 ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -30 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -2 $out/smoke.log
( time timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err ); tail -c 3000 $out/bench_default.json
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; ( env "$@" timeout 300 $B 2> $out/bench_$name.err | tail -1 > $out/bench_$name.json ); echo "$name rc=$?"; }
run ring1 RTM_RING2=1
run ring1_fwd0 RTM_RING2=1 RTM_RING2_FWD=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c7/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'], 'e2e', d.get('e2e') and round(d['e2e']['value']), d.get('parity_checked'))
    except Exception as e:
        print(f, 'ERR', e)
PY

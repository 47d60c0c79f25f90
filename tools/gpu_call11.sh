#!/bin/bash
# round 2, GPU call 11: ring_kernel for the adaptive operator / single backward steps: parity + A/B on C3, C5 adaptive, C5 taylor R=8/12
out=gpurun_out/c11; mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -8 $out/pytest_gpu.log
run() { name=$1; cfg=$2; shift; shift; ( env "$@" timeout 600 python bench.py --config $cfg --warmup 1 --no-cpu-baseline --no-e2e 2> $out/$name.err | tail -1 > $out/$name.json ); echo "$name rc=$?"; }
for cfg in c3 c5 c5:4 c5:8:taylor c5:12:taylor; do
  n=$(echo $cfg | tr ':' '_')
  run ${n}_ring1 $cfg RTM_RING2=1
  run ${n}_ring0 $cfg RTM_RING2=0
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c11/c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $out/memcheck_small.log 2>&1
tail -3 $out/memcheck_small.log

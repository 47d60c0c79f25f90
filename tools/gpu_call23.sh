#!/bin/bash
# round 2, GPU call 23: ring_kernel next to SINGLE backward steps (Taylor radius 8 / 12): off / on / on with a high-priority ring stream;
# driver tests after the host-buffer change
out=gpurun_out/c23; mkdir -p $out
run() { name=$1; shift; ( timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err ); echo "$name rc=$?"; }
for R in 8 12; do
  run t${R}_on --config c5:$R:taylor --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
  RTM_RING_PRIO=1 run t${R}_on_prio --config c5:$R:taylor --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
  RTM_RING2_BWD=0 run t${R}_off --config c5:$R:taylor --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
done
( timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -q -x --timeout 600 ) > $out/pytest_driver.log 2>&1; tail -2 $out/pytest_driver.log
python tools/run_driver_job.py --config c2 --gpus 1 > $out/driver_c2_1gpu.json 2> $out/driver_c2.err; tail -c 500 $out/driver_c2_1gpu.json; echo
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c23/t*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline')
        print(f.split('/')[-1], d['config']['workload'], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), d.get('gpu_launches'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY

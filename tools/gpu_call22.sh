#!/bin/bash
# round 2, GPU call 22: (a) Taylor radius 8 line re-measured (one slow sample in call 21), with and without the mbarrier suspend hint;
# (b) same-bin fast path of the adaptive streaming kernel (variant ls4same) against the default build
out=gpurun_out/c22; mkdir -p $out
run() { name=$1; shift; ( timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err ); echo "$name rc=$?"; }
H=rtm_gpu_b200/variants/librtm_b200_hint0.so
V=rtm_gpu_b200/variants/librtm_b200_ls4same.so
for i in 1 2 3; do run t8_base_$i --config c5:8:taylor --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda; done
for i in 1 2; do RTM_LIB_PATH=$H run t8_hint0_$i --config c5:8:taylor --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda; done
RTM_RING2_BWD=0 run t8_ring2bwd0 --config c5:8:taylor --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
RTM_RING2=0 run t8_ring2_0 --config c5:8:taylor --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
( RTM_LIB_PATH=$V timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -q -x --timeout 600 -k adaptive ) > $out/pytest_ls4same.log 2>&1; tail -2 $out/pytest_ls4same.log
run r4_base --config c5:4 --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
RTM_LIB_PATH=$V run r4_ls4same --config c5:4 --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
run c4_base --config c4 --nt 1000 --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
RTM_LIB_PATH=$V run c4_ls4same --config c4 --nt 1000 --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c22/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline')
        print(f.split('/')[-1], d['config']['workload'], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), d.get('gpu_launches'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY

#!/bin/bash
# round 2, GPU call 18: adaptive tile kernels with the warp-uniform term loop (variant build lsuni) against the default build
out=gpurun_out/c18; mkdir -p $out
V=rtm_gpu_b200/variants/librtm_b200_lsuni.so
( RTM_LIB_PATH=$V timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_vs_ref_cuda.py -m gpu -q -x --timeout 600 ) > $out/pytest_lsuni.log 2>&1; tail -2 $out/pytest_lsuni.log
run() { name=$1; shift; ( timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err ); echo "$name rc=$?"; }
for cfg in c5:8 c5:12 c5 c5:6 c3; do
  n=$(echo $cfg | tr ':' '_')
  run ${n}_base --config $cfg --warmup 1 --no-cpu-baseline --no-ref-cuda
  RTM_LIB_PATH=$V run ${n}_lsuni --config $cfg --warmup 1 --no-cpu-baseline --no-ref-cuda
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c18/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline')
        print(f.split('/')[-1], d['config']['workload'], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY

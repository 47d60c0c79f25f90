#!/bin/bash
# round 2, GPU call 17: adaptive kernels after (a) the double path of the first stencil term behind a real branch, (b) uniform
# term loop + selects in the streaming form; ring launches next to ib / thin by default.  Parity first, then bench lines.
out=gpurun_out/c17; mkdir -p $out
( time timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_driver.py tests/test_gpu_named_configs.py -m gpu -q -x --timeout 900 ) > $out/pytest.log 2>&1
echo "rc=$?" >> $out/pytest.log; tail -6 $out/pytest.log
run() { name=$1; shift; ( timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err ); echo "$name rc=$?"; }
run c5_r4 --config c5:4 --warmup 1 --no-cpu-baseline --no-ref-cuda
run c4_nt2000 --config c4 --nt 2000 --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
run c5_adaptive --config c5 --warmup 1 --no-cpu-baseline
run c5_r8 --config c5:8 --warmup 1 --no-cpu-baseline --no-ref-cuda
run c5_r12 --config c5:12 --warmup 1 --no-cpu-baseline --no-ref-cuda
run c3 --config c3 --warmup 1 --no-cpu-baseline
run c3_b120 --config c3 --warmup 1 --shots-per-step 120 --steps 2 --no-cpu-baseline --no-ref-cuda
run default_nt301 --nt 301 --steps 3 --warmup 1 --no-cpu-baseline --no-ref-cuda
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c17/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline')
        print(f.split('/')[-1], d['config']['workload'], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'parity', d.get('parity_checked'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
du -sh gpurun_out

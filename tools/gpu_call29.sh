#!/bin/bash
# round 2, GPU call 29: defaults changed (segments of 12 tiles, forward look-ahead 296 from 48 shots per launch): parity + bench lines
out=gpurun_out/c29; mkdir -p $out
( timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_fullsize.py "tests/test_gpu_named_configs.py::test_c2_whole_shot_full_time_axis" "tests/test_gpu_named_configs.py::test_c4_shape_20000_wide_adaptive" -m gpu -q -x --timeout 600 ) > $out/pytest.log 2>&1; tail -2 $out/pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -2 $out/smoke.log
python bench.py --nt 301 --steps 3 --warmup 1 --no-cpu-baseline --no-ref-cuda > $out/default_nt301.json 2>/dev/null
python bench.py --config c4 --nt 400 --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > $out/c4_nt400.json 2>/dev/null
( time timeout 1200 python bench.py > $out/bench_default.json 2> $out/bench_default.err )
python - <<'PY'
import json
for n in ('default_nt301','c4_nt400','bench_default'):
    d=json.loads(open('gpurun_out/c29/'+n+'.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(n, round(d['value']), 'bwd', round(1e3*r['avg_launch_ms'],1), 'fwd', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'], 'e2e', d.get('e2e') and round(d['e2e']['value']), 'parity', d.get('parity_checked'), 'dram', (r.get('dram') or {}).get('frac'))
PY

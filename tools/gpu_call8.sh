#!/bin/bash
# round 2, GPU call 8: mbarrier waits with a suspend hint; the default bench under the power cap, ring2_fwd on / off
out=gpurun_out/c8; mkdir -p $out
( time timeout 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py tests/test_gpu_parity.py tests/test_gpu_named_configs.py -m gpu -q --timeout 600 --durations=6 ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -14 $out/pytest_gpu.log
( time timeout 900 python bench.py --no-cpu-baseline > $out/bench_default_fwd1.json 2> $out/bench_default_fwd1.err )
( time RTM_RING2_FWD=0 timeout 900 python bench.py --no-cpu-baseline > $out/bench_default_fwd0.json 2> $out/bench_default_fwd0.err )
( time RTM_STREAM2=0 RTM_RING2=0 timeout 900 python bench.py --no-cpu-baseline > $out/bench_default_r1form.json 2> $out/bench_default_r1form.err )
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c8/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'], 'e2e', d.get('e2e') and round(d['e2e']['value']), 'dram', r.get('dram') and round(r['dram']['frac'],3))
    except Exception as e:
        print(f, 'ERR', e)
PY
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream2 -s 2 -c 1 -o $out/prof_stream_bwd $P > $out/ncu_full.log 2>&1

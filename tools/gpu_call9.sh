#!/bin/bash
# round 2, GPU call 9: single-step streaming forward kernel: parity, A/B at boost clocks and under the power cap
out=gpurun_out/c9; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_vs_ref_cuda.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -x ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -8 $out/pytest_gpu.log
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; ( env "$@" timeout 300 $B 2> $out/bench_$name.err | tail -1 > $out/bench_$name.json ); echo "$name rc=$?"; }
run s1_on RTM_STREAM1_FWD=1
run s1_off RTM_STREAM1_FWD=0
run s1_on_seg4 RTM_STREAM1_FWD=1 RTM_SEG_TILES=4
run s1_on_seg16 RTM_STREAM1_FWD=1 RTM_SEG_TILES=16
( time timeout 900 python bench.py --no-cpu-baseline > $out/bench_default_s1.json 2> $out/bench_default_s1.err )
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c9/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'e2e', d.get('e2e') and round(d['e2e']['value']), d.get('parity_checked'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream1 -s 4 -c 1 -o $out/prof_stream1 $P > $out/ncu_full.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_stream.py > $out/memcheck.log 2>&1
tail -3 $out/memcheck.log

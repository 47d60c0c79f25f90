#!/bin/bash
# round 2, GPU call 5: thin-frame slot k from the streamed halo: bit-exactness, A/B timing, profiles
out=gpurun_out/c5; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py tests/test_gpu_shapes.py -q --timeout 300 -x ) > $out/pytest_stream.log 2>&1
echo "rc=$?" >> $out/pytest_stream.log
tail -15 $out/pytest_stream.log
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; ( env "$@" timeout 300 $B 2> $out/bench_$name.err | tail -1 > $out/bench_$name.json ); echo "$name rc=$?"; }
run tile RTM_STREAM2=0
run s8 RTM_STREAM2=1 RTM_FUSE2_FWD=0 RTM_SEG_TILES=8
run s5 RTM_STREAM2=1 RTM_FUSE2_FWD=0 RTM_SEG_TILES=5
run s12 RTM_STREAM2=1 RTM_FUSE2_FWD=0 RTM_SEG_TILES=12
run s8f RTM_STREAM2=1 RTM_FUSE2_FWD=1 RTM_SEG_TILES=8
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c5/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), 'exec frac', round(r['executed']['frac'],3))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e"
RTM_FUSE2_FWD=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
RTM_FUSE2_FWD=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream2 -s 2 -c 2 -o $out/prof_stream_bwd $P > $out/ncu_full.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_stream.py > $out/memcheck.log 2>&1
tail -4 $out/memcheck.log

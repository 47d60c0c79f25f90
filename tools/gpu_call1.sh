#!/bin/bash
# round 2, GPU call 1: correctness of the streaming kernels, A/B timing, named-config parity, profiles
out=gpurun_out/c1; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,memory.total --format=csv > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt; free -g >> $out/gpu.txt
( time timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py -q --timeout 300 ) > $out/pytest_stream.log 2>&1
echo "rc=$?" >> $out/pytest_stream.log
tail -5 $out/pytest_stream.log
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; ( env "$@" timeout 300 $B 2> $out/bench_$name.err | tail -1 > $out/bench_$name.json ); echo "$name rc=$?"; }
run tile RTM_STREAM2=0
run s8 RTM_STREAM2=1 RTM_FUSE2_FWD=0 RTM_SEG_TILES=8
run s4 RTM_STREAM2=1 RTM_FUSE2_FWD=0 RTM_SEG_TILES=4
run s16 RTM_STREAM2=1 RTM_FUSE2_FWD=0 RTM_SEG_TILES=16
run s8f RTM_STREAM2=1 RTM_FUSE2_FWD=1 RTM_SEG_TILES=8
run s16f RTM_STREAM2=1 RTM_FUSE2_FWD=1 RTM_SEG_TILES=16
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c1/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
( time timeout 1500 python -m pytest tests/test_gpu_named_configs.py -q --timeout 900 ) > $out/pytest_named.log 2>&1
echo "rc=$?" >> $out/pytest_named.log
tail -8 $out/pytest_named.log
P="python bench.py --nt 41 --steps 1 --warmup 0 --shots-per-step 8 --no-cpu-baseline --no-e2e"
RTM_FUSE2_FWD=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
RTM_FUSE2_FWD=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream2 -s 8 -c 6 -o $out/prof_stream $P > $out/ncu_full.log 2>&1
ls -la $out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_stream.py > $out/memcheck.log 2>&1
tail -5 $out/memcheck.log

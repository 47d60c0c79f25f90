#!/bin/bash
# round 2, GPU call 12: ring kernel without coefficient staging (37.6 KB), streaming kernel capped at 88 registers (thin-frame CTAs co-reside),
# batch size
out=gpurun_out/c12; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_vs_ref_cuda.py -m gpu -q --timeout 600 -x ) > $out/pytest_gpu.log 2>&1
echo "rc=$?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; ( env "$@" timeout 300 $B 2> $out/bench_$name.err | tail -1 > $out/bench_$name.json ); echo "$name rc=$?"; }
run reg88 A=1
run reg96 RTM_LIB_PATH=$PWD/rtm_gpu_b200/librtm_b200_reg96.so
run reg88_b RTM_LIB_PATH=$PWD/rtm_gpu_b200/librtm_b200.so
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --shots-per-step 48"
run reg88_s48 A=1
B="python bench.py --nt 301 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --shots-per-step 64"
run reg88_s64 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c12/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'bwd us/shot', round(1e3*r['avg_launch_ms']/d['config']['shots_per_step_per_gpu'],3), 'fwd us/shot', round(1e3*r['forward_step']['avg_launch_ms']/d['config']['shots_per_step_per_gpu'],3), d['clocks'].get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
( time timeout 900 python bench.py --no-cpu-baseline > $out/bench_default.json 2> $out/bench_default.err )
python -c "
import json;d=json.loads(open('gpurun_out/c12/bench_default.json').read().strip().splitlines()[-1]);r=d['roofline'];print('default',round(d['value']),'bwd',round(1e3*r['avg_launch_ms'],1),'fwd',round(1e3*r['forward_step']['avg_launch_ms'],1),d['clocks'],'e2e',round(d['e2e']['value']))"

#!/bin/bash
# round 2, GPU call 15: adaptive operator in the streaming form (parity, sanitizers, bench lines) + racecheck with the
# named-barrier checking build + the records of the default bench
out=gpurun_out/c15; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py tests/test_gpu_parity.py tests/test_gpu_vs_ref_cuda.py "tests/test_gpu_named_configs.py::test_c4_shape_20000_wide_adaptive" -m gpu -q -x --timeout 600 ) > $out/pytest_ls.log 2>&1
echo "rc=$?" >> $out/pytest_ls.log; tail -8 $out/pytest_ls.log
S="compute-sanitizer --racecheck-report analysis --print-limit 12"
RTM_LIB_PATH=rtm_gpu_b200/variants/librtm_b200_barsync.so timeout 600 $S --tool racecheck python tools/sanitize_stream.py > $out/racecheck_barsync.log 2>&1; tail -2 $out/racecheck_barsync.log
RTM_SAN_LS=1 RTM_LIB_PATH=rtm_gpu_b200/variants/librtm_b200_barsync.so timeout 600 $S --tool racecheck python tools/sanitize_stream.py > $out/racecheck_barsync_ls.log 2>&1; tail -2 $out/racecheck_barsync_ls.log
timeout 600 $S --tool racecheck python tools/sanitize_stream.py > $out/racecheck_mbarrier.log 2>&1; tail -1 $out/racecheck_mbarrier.log
RTM_SAN_LS=1 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_stream.py > $out/memcheck_ls.log 2>&1; tail -2 $out/memcheck_ls.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_stream.py > $out/memcheck.log 2>&1; tail -2 $out/memcheck.log
run() { name=$1; shift; ( time timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err ); echo "$name rc=$?"; }
run c5_r4 --config c5:4 --warmup 1 --no-cpu-baseline
RTM_FUSE2=0 run c5_r4_single --config c5:4 --warmup 1 --no-cpu-baseline --no-ref-cuda
run c4 --config c4 --warmup 1 --steps 2 --no-cpu-baseline
RTM_FUSE2=0 run c4_single --config c4 --warmup 1 --steps 1 --no-cpu-baseline --no-ref-cuda
run bench_default
run bench_reference --impl reference --steps 2 --warmup 1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c15/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline')
        if not r: print(f.split('/')[-1], d.get('value'), d.get('impl')); continue
        print(f.split('/')[-1], d['config']['workload'], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'frac', round(r['frac'],3), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'e2e', d.get('e2e') and round(d['e2e']['value']), 'parity', d.get('parity_checked'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
du -sh gpurun_out

#!/bin/bash
# round 2, GPU call 16: adaptive streaming kernel after hoisting the per-cell lookups (bench + ncu), ring launches next to ib / thin (A/B)
out=gpurun_out/c16; mkdir -p $out
( timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -q -x --timeout 600 ) > $out/pytest_stream.log 2>&1; tail -2 $out/pytest_stream.log
run() { name=$1; shift; ( timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err ); echo "$name rc=$?"; }
run c5_r4 --config c5:4 --warmup 1 --no-cpu-baseline --no-ref-cuda
run c4_nt2000 --config c4 --nt 2000 --warmup 1 --steps 2 --no-cpu-baseline --no-ref-cuda
A="--nt 301 --steps 3 --warmup 1 --no-cpu-baseline --no-ref-cuda"
RTM_RING_PAR=0 run ab_par0 $A
RTM_RING_PAR=1 run ab_par1 $A
RTM_RING_PAR=1 RTM_RING_PRIO=1 run ab_par1_prio $A
RTM_RING_PAR=0 run ab_par0_b $A
RTM_RING_PAR=1 RTM_RING_PRIO=1 run ab_par1_prio_b $A
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c16/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline')
        print(f.split('/')[-1], d['config']['workload'], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
P="python bench.py --config c5:4 --nt 25 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-ref-cuda"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream2 -s 2 -c 1 -o $out/prof_stream_ls $P > $out/ncu_full_ls.log 2>&1
python tools/ncu_summary.py full $out/prof_stream_ls.ncu-rep > $out/stream_ls_full.txt 2>&1
ncu -i $out/prof_stream_ls.ncu-rep --page source --csv 2>/dev/null | gzip > $out/stream_ls_source.csv.gz
P="python bench.py --config c5:4:taylor --nt 25 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-ref-cuda"
timeout 600 ncu --set full --clock-control none -k regex:stream2 -s 2 -c 1 -o $out/prof_stream_te $P > $out/ncu_full_te.log 2>&1
python tools/ncu_summary.py full $out/prof_stream_te.ncu-rep > $out/stream_te_full.txt 2>&1
rm -f $out/prof_stream_te.ncu-rep
ls -la $out; du -sh gpurun_out

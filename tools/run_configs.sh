#!/bin/bash
# One bench line per BASELINE.json configuration other than the headline (C3, C5 radius sweep, C4), into gpurun_out/cfg/.
out=gpurun_out/cfg; mkdir -p $out
run() { name=$1; shift; ( time timeout 1500 python bench.py "$@" > $out/$name.json 2> $out/$name.err ); echo "$name rc=$? $(tail -c 300 $out/$name.json | head -c 0)"; }
run c3 --config c3 --warmup 1
run c3_b120 --config c3 --warmup 1 --shots-per-step 120 --steps 2 --no-cpu-baseline
for R in 4 5 6 7 8 9 10 11 12; do run c5_r$R --config c5:$R --warmup 1 --no-cpu-baseline; done
run c5_adaptive --config c5 --warmup 1
for R in 4 8 12; do run c5_taylor_r$R --config c5:$R:taylor --warmup 1 --no-cpu-baseline; done
run c4 --config c4 --warmup 1 --steps 2 --no-cpu-baseline
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/cfg/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], d['config']['workload'], round(d['value']), 'bwd us', round(1e3*r['avg_launch_ms'],1), 'frac', round(r['frac'],3), 'fwd us', round(1e3*r['forward_step']['avg_launch_ms'],1), 'frac', round(r['forward_step']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'e2e', d.get('e2e') and round(d['e2e']['value']), 'parity', d.get('parity_checked'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY

#!/bin/bash
# tools/final_check.sh -- one gpurun call: GPU test suite, smoke, both bench arms, batch-size sweep.
# Everything lands in gpurun_out/final/.
out=gpurun_out/final; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
timeout 400 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err
for b in 8 12 16 24 32; do
  timeout 300 python bench.py --nt 301 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --shots-per-step $b 2>/dev/null | tail -1 > $out/sweep_b$b.json
done
# timing-only variants (wrong results by design): cost of the ring tiles / of their one-way phase
tools/sweep_lib.sh base noring nooneway > $out/sweep_ring.txt 2>&1
cat $out/sweep_ring.txt
tail -3 $out/pytest_gpu.log; cat $out/smoke.log | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline',{})
        print(f.split('/')[-1], d.get('value'), d.get('config',{}).get('shots_per_step_per_gpu'), 'bwd us', 1e3*r.get('avg_launch_ms',0), 'fwd us', 1e3*r.get('forward_step',{}).get('avg_launch_ms',0), 'e2e', d.get('e2e',{}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY

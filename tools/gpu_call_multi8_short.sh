#!/bin/bash
# round 2, final 8-GPU check (short: 8x charge): in-process path (thread per GPU + one ncclReduce) next to torchrun, multi-GPU tests
out=gpurun_out/multi8f; mkdir -p $out
( timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 150 ) > $out/pytest_multi.log 2>&1; tail -2 $out/pytest_multi.log
( timeout 200 python bench.py --inproc --gpus 8 --nt 1501 --steps 2 --warmup 1 > $out/bench_inproc.json 2> $out/bench_inproc.err )
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --nt 1501 --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > $out/bench_torchrun.json 2> $out/bench_torchrun.err )
python - <<PY
import json
for n in ('bench_inproc','bench_torchrun'):
    try:
        d=json.loads(open('$out/'+n+'.json').read().strip().splitlines()[-1])
        print(n, 'value', round(d['value']), 'per gpu', round(d['per_gpu_value']), 'n_gpus', d['n_gpus'], d.get('reduce_backend'), d.get('reduce_ms'))
    except Exception as e: print(n,'ERR',e)
PY

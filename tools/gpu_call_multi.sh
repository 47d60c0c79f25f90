#!/bin/bash
# round 2, multi-GPU call (N = $1): the north_star path -- one host thread per GPU in ONE process, in-process ncclReduce --
# next to the torchrun path the driver's SCALE run measures
N=${1:-2}
out=gpurun_out/multi$N; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/gpus.txt
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 ) > $out/pytest_multi.log 2>&1; tail -4 $out/pytest_multi.log
NT=${2:-1501}
( time timeout 600 python bench.py --inproc --gpus $N --nt $NT --steps 2 --warmup 1 > $out/bench_inproc.json 2> $out/bench_inproc.err ); tail -c 700 $out/bench_inproc.json; echo
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --nt $NT --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_torchrun.json 2> $out/bench_torchrun.err ); tail -c 400 $out/bench_torchrun.json; echo
( time timeout 900 python tools/run_driver_job.py --config c2 --gpus $N --shots 64 > $out/driver_c2.json 2> $out/driver_c2.err ); cat $out/driver_c2.json
( time timeout 900 python tools/run_driver_job.py --config c3 --gpus $N --shots 240 > $out/driver_c3.json 2> $out/driver_c3.err ); cat $out/driver_c3.json
python - <<PY
import json
for n in ('bench_inproc','bench_torchrun'):
    try:
        d=json.loads(open('$out/'+n+'.json').read().strip().splitlines()[-1])
        print(n, 'value', round(d['value']), 'per gpu', round(d['per_gpu_value']), 'n_gpus', d['n_gpus'], d.get('reduce_backend'), d.get('reduce_ms'))
    except Exception as e: print(n,'ERR',e)
PY

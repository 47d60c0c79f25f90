#!/bin/bash
# round 2, GPU call 33 (last GPU seconds): the region construction now lives in make_stream_regions() -- streaming tests + smoke
out=gpurun_out/c33; mkdir -p $out
( timeout 80 python -m pytest tests/test_gpu_stream.py -m gpu -q --timeout 80 ) > $out/pytest_stream.log 2>&1; echo "rc=$?" >> $out/pytest_stream.log; tail -3 $out/pytest_stream.log
timeout 40 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -2 $out/smoke.log

#!/bin/bash
# round 2, GPU call 31: the streaming tests with the pair policy by streamed cells (the one test call 30 ran against the previous build)
out=gpurun_out/c31; mkdir -p $out
( timeout 200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fuse2.py -m gpu -q --timeout 200 ) > $out/pytest_stream.log 2>&1; echo "rc=$?" >> $out/pytest_stream.log; tail -4 $out/pytest_stream.log

#!/bin/bash
# round 2, GPU call 30: parity files after the change of defaults (segment length 12, pair policy by streamed cells)
out=gpurun_out/c30; mkdir -p $out
( time timeout 600 python -m pytest tests -m gpu -q --timeout 600 --deselect "tests/test_gpu_named_configs.py::test_c5_shape_4096_wide" ) > $out/pytest_gpu.log 2>&1; echo "rc=$?" >> $out/pytest_gpu.log; tail -5 $out/pytest_gpu.log

#!/bin/bash
# round 2, GPU call 32 (last GPU-minutes): tests whose launches depend on the automatic pair policy, with the final build
out=gpurun_out/c32; mkdir -p $out
( timeout 140 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_driver.py "tests/test_gpu_named_configs.py::test_c4_shape_20000_wide_adaptive" "tests/test_gpu_named_configs.py::test_c2_whole_shot_full_time_axis" -m gpu -q --timeout 130 ) > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log

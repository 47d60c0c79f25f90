#!/bin/bash
# round 2, GPU call 21: every configuration's bench line with the final code, the drop-in executable on C2 / C3, and the
# full-size 20000 x 5000 parity case (adaptive operator through the streaming kernels) against the reference CUDA build
bash tools/run_configs.sh > gpurun_out/run_configs.log 2>&1; tail -22 gpurun_out/run_configs.log
out=gpurun_out/c21; mkdir -p $out
python tools/run_driver_job.py --config c3 --gpus 1 > $out/driver_c3_1gpu.json 2> $out/driver_c3.err; tail -c 700 $out/driver_c3_1gpu.json; echo
python tools/run_driver_job.py --config c2 --gpus 1 > $out/driver_c2_1gpu.json 2> $out/driver_c2.err; tail -c 700 $out/driver_c2_1gpu.json; echo
( time RTM_TEST_SLOW=1 timeout 1500 python -m pytest "tests/test_gpu_named_configs.py::test_c4_shape_20000_wide_adaptive" -m gpu -q --timeout 1400 ) > $out/pytest_c4_full.log 2>&1; tail -6 $out/pytest_c4_full.log
du -sh gpurun_out

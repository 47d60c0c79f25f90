#!/bin/bash
# round 2, GPU call 10: one bench line per BASELINE configuration (C3, C5 sweep, C4) + the drop-in executable on C2 / C3 (1 GPU)
bash tools/run_configs.sh
out=gpurun_out/cfg
( time timeout 900 python tools/run_driver_job.py --config c2 --gpus 1 --shots 64 > $out/driver_c2_1gpu.json 2> $out/driver_c2_1gpu.err ); cat $out/driver_c2_1gpu.json
( time timeout 900 python tools/run_driver_job.py --config c3 --gpus 1 --shots 240 > $out/driver_c3_1gpu.json 2> $out/driver_c3_1gpu.err ); cat $out/driver_c3_1gpu.json

#!/bin/bash
# tools/final_records.sh -- one gpurun call: the default bench line, the ncu launch list and full captures of the step kernels
out=gpurun_out/records; mkdir -p $out
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -c 600 $out/bench.json
B="python bench.py --nt 81 --steps 1 --warmup 0 --shots-per-step 8 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file $out/launches.csv $B > $out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwd_step_kernel -s 10 -c 1 -f -o $out/fwd $B > $out/fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_step_kernel -s 6 -c 1 -f -o $out/bwdring $B > $out/bwdring.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd2_step_kernel -s 4 -c 2 -f -o $out/bwd2 $B > $out/bwd2.log 2>&1
ls -la $out

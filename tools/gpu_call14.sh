#!/bin/bash
# round 2, GPU call 14: racecheck attribution of the round-2 kernels + the records call 13 could not bring back
out=gpurun_out/c14; mkdir -p $out
# (a) streaming pair kernels + thin frame + ring kernel; (b) single steps only: ring_kernel + tile kernels
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 60 python tools/sanitize_stream.py > $out/racecheck_stream.log 2>&1; tail -2 $out/racecheck_stream.log
RTM_FUSE2=0 RTM_FUSE2_FWD=0 timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 60 python tools/sanitize_stream.py > $out/racecheck_single.log 2>&1; tail -2 $out/racecheck_single.log
RTM_RING2=0 timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 60 python tools/sanitize_stream.py > $out/racecheck_noring2.log 2>&1; tail -2 $out/racecheck_noring2.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_stream.py > $out/memcheck.log 2>&1; tail -2 $out/memcheck.log
grep -h "hazard detected\|Race reported" $out/racecheck_stream.log | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -20
( time timeout 1200 python bench.py > $out/bench_default.json 2> $out/bench_default.err )
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err )
P="python bench.py --nt 25 --steps 1 --warmup 0 --shots-per-step 32 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/launches.csv $P > $out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream2 -s 2 -c 2 -o $out/prof_stream_bwd $P > $out/ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ring_kernel -s 4 -c 2 -o $out/prof_ring $P > $out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fwd_step -s 4 -c 1 -o $out/prof_fwd $P > $out/ncu_full3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:thin_frame -s 2 -c 1 -o $out/prof_thin $P > $out/ncu_full4.log 2>&1
ls -la $out; du -sh gpurun_out

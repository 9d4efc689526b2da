#!/bin/bash
# round-end measurement set (one B200): bench lines, ncu launch lists, ncu --set full
TAG=${TAG:-r1o}
O=gpurun_out
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 200 python bench.py --steps 50 --warmup 5 --dtype f32 > $O/${TAG}_bench_f32.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>> $O/${TAG}_bench.err
timeout 300 ncu --metrics $M --clock-control none -s 42 -c 14 --csv --log-file $O/${TAG}_launches.csv python scripts/one_step.py 256 3 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 56 -c 14 --csv --log-file $O/${TAG}_launches_T32.csv python scripts/one_step.py 32 4 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 42 -c 14 --csv --log-file $O/${TAG}_launches_f32.csv python scripts/one_step.py 256 3 --f32 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lauum_grad -s 2 -c 1 -f -o $O/${TAG}_lauum python scripts/one_step.py 256 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 23 -c 2 -f -o $O/${TAG}_step python scripts/one_step.py 256 2 > /dev/null 2>&1
timeout 600 python scripts/configs.py > $O/${TAG}_configs_c4_c5.jsonl 2> $O/${TAG}_configs.err
ls -la $O | tail -20
tail -c 600 $O/${TAG}_bench.json

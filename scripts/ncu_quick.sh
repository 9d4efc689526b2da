#!/bin/bash
# launch list + full capture of lauum and two k_step launches (fp64, T=256)
TAG=${TAG:-r1l}
O=gpurun_out
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 ncu --metrics $M --clock-control none -s 42 -c 14 --csv --log-file $O/${TAG}_launches.csv python scripts/one_step.py 256 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lauum_grad -s 2 -c 1 -f -o $O/${TAG}_lauum python scripts/one_step.py 256 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 23 -c 2 -f -o $O/${TAG}_step python scripts/one_step.py 256 2 > /dev/null 2>&1
ls -la $O | grep $TAG

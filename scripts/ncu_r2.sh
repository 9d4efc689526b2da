#!/bin/bash
# round-2 profile set: launch list of one hb_nll_grad_batched call (4 launches:
# k_prep, k_fused, k_task_final, k_reduce_final) + one --set full capture of the
# persistent kernel (fp64, 256 x 512 x 8)
TAG=${TAG:-r2}
O=gpurun_out
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 ncu --metrics $M --clock-control none -s 12 -c 4 --csv --log-file $O/${TAG}_launches.csv python scripts/one_step.py 256 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -f -o $O/${TAG}_fused python scripts/one_step.py 256 2 > /dev/null 2>&1
ls -la $O | grep $TAG

#!/bin/bash
# final build: per-launch ncu metrics of all 62 conv launches of one pass (no cache flush between kernels)
cd /root/repo
timeout 100 python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/one_pass_plain_r2.log 2>&1 && \
AYQ_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size --cache-control none --clock-control none -k regex:conv_tma -s 62 -c 62 --csv --page raw --log-file gpurun_out/conv_tma_full_r2.csv python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/ncu_full_r2.log 2>&1
wc -l gpurun_out/conv_tma_full_r2.csv; tail -2 gpurun_out/ncu_full_r2.log

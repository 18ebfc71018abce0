#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "past_its or batch_64" > gpurun_out/pytest_s13.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s13.txt
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_plain_r2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 201 -c 134 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_r2.log 2>&1
timeout 100 python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/one_pass_plain_r2.log 2>&1 && \
AYQ_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size --cache-control none --clock-control none -k regex:conv_tma -s 62 -c 62 --csv --page raw --log-file gpurun_out/conv_tma_full_r2.csv python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/ncu_full_r2.log 2>&1
tail -3 gpurun_out/pytest_s13.txt; wc -l gpurun_out/launches_r2.csv gpurun_out/conv_tma_full_r2.csv

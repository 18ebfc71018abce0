#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or host_entry or different_streams or nms_corner or full_size or production_plan or fresh" > gpurun_out/pytest_s4.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s4.txt
AYQ_PLAN_DUMP=1 timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp4_base.txt 2>&1
AYQ_HALO_MIN_NP=2 timeout 200 python tools/exp.py --tag np2 --ops > gpurun_out/exp4_np2.txt 2>&1
AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/role4.txt 2>&1
timeout 100 python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/one_pass_plain4.txt 2>&1 && AYQ_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_tma -s 62 -c 62 --csv --page raw --log-file gpurun_out/conv_tma_l2_s4.csv python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/ncu_s4.log 2>&1
tail -n 3 gpurun_out/exp4_*.txt; tail -n 5 gpurun_out/pytest_s4.txt

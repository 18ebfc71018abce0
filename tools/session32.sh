#!/bin/bash
cd /root/repo
AYQ_PLAN_DUMP=1 timeout 200 python tools/exp.py --tag tuned --ops > gpurun_out/exp32_tuned.txt 2>&1
AYQ_AUTOTUNE=0 timeout 200 python tools/exp.py --tag untuned --ops > gpurun_out/exp32_untuned.txt 2>&1
timeout 200 python tools/exp.py --tag tuned2 > gpurun_out/exp32_tuned2.txt 2>&1
grep -h "images/s" gpurun_out/exp32_*.txt; grep "^tune" gpurun_out/exp32_tuned.txt | grep -c "one$"; grep "^tune" gpurun_out/exp32_tuned.txt | grep "one$"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_s32.txt 2>&1; tail -3 gpurun_out/pytest_s32.txt

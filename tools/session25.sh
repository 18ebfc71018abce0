#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_s25.txt 2>&1; tail -3 gpurun_out/pytest_s25.txt
timeout 200 python tools/exp.py --tag p1persist --ops > gpurun_out/exp25.txt 2>&1; grep -h "images/s\|sum of" gpurun_out/exp25.txt; grep -o "absmax=[0-9.]* Conv_P1=[0-9.]*" gpurun_out/exp25.txt
AYQ_P1_NO_PERSIST=1 timeout 200 python tools/exp.py --tag p1plain --ops > gpurun_out/exp25b.txt 2>&1; grep -h "images/s\|sum of" gpurun_out/exp25b.txt; grep -o "absmax=[0-9.]* Conv_P1=[0-9.]*" gpurun_out/exp25b.txt

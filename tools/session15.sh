#!/bin/bash
cd /root/repo
AYQ_P1_FUSE=1 AYQ_P1_FUSE_D=0 timeout 200 python tools/exp.py --tag fuse_d0 --ops > gpurun_out/exp15_d0.txt 2>&1
AYQ_P1_FUSE=1 AYQ_P1_FUSE_D=1 timeout 200 python tools/exp.py --tag fuse_d1 --ops > gpurun_out/exp15_d1.txt 2>&1
AYQ_P1_FUSE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "all_golden and tma or host_entry or full_size" > gpurun_out/pytest_s15.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s15.txt
grep -h "images/s\|sum of" gpurun_out/exp15_*.txt; tail -n 3 gpurun_out/pytest_s15.txt; grep -ho "absmax=[0-9.]* Conv_P1=[0-9.]*" gpurun_out/exp15_*.txt

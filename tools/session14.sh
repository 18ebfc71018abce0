#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or host_entry or fresh or past_its" > gpurun_out/pytest_s14.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s14.txt
timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp14_base.txt 2>&1
grep -h "images/s\|sum of" gpurun_out/exp14_*.txt; tail -n 3 gpurun_out/pytest_s14.txt; grep -o "absmax=[0-9.]* Conv_P1=[0-9.]*" gpurun_out/exp14_base.txt

#!/bin/bash
cd /root/repo
timeout 100 ./tools/ubench/pipes > gpurun_out/pipes_r2.txt 2>&1
timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp7_base.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tma or full_size or production" > gpurun_out/pytest_s7.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s7.txt
grep -h "images/s\|sum of" gpurun_out/exp7_*.txt; tail -n 3 gpurun_out/pytest_s7.txt; cat gpurun_out/pipes_r2.txt

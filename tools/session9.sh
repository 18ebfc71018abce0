#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_s9.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s9.txt
timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp9_base.txt 2>&1
timeout 200 python tools/exp.py --tag k6 --k 6 --ops > gpurun_out/exp9_k6.txt 2>&1
timeout 200 python tools/exp.py --tag k4 --k 4 > gpurun_out/exp9_k4.txt 2>&1
timeout 200 python tools/exp.py --tag b512 --batch 512 > gpurun_out/exp9_b512.txt 2>&1
grep -h "images/s\|sum of" gpurun_out/exp9_*.txt; tail -n 3 gpurun_out/pytest_s9.txt

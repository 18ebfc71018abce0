#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_s6.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s6.txt
timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp6_base.txt 2>&1
timeout 200 python tools/exp.py --tag k6 --k 6 > gpurun_out/exp6_k6.txt 2>&1
timeout 200 python tools/exp.py --tag b512 --batch 512 > gpurun_out/exp6_b512.txt 2>&1
AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/role6.txt 2>&1
grep -h "images/s\|sum of" gpurun_out/exp6_*.txt; tail -n 5 gpurun_out/pytest_s6.txt

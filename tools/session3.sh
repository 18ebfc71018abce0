#!/bin/bash
cd /root/repo
timeout 120 ./tools/ubench/ubench > gpurun_out/ubench_r2.txt 2>&1; echo "ubench rc=$?" >> gpurun_out/ubench_r2.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_s3.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s3.txt
timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp3_base.txt 2>&1
AYQ_NO_P1_FUSE=1 timeout 200 python tools/exp.py --tag nofuse --ops > gpurun_out/exp3_nofuse.txt 2>&1
timeout 200 python tools/exp.py --tag b512 --batch 512 > gpurun_out/exp3_b512.txt 2>&1
timeout 200 python tools/exp.py --tag k6 --k 6 --ops > gpurun_out/exp3_k6.txt 2>&1
tail -n 3 gpurun_out/exp3_*.txt; tail -n 30 gpurun_out/pytest_s3.txt; cat gpurun_out/ubench_r2.txt

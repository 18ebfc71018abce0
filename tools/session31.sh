#!/bin/bash
cd /root/repo
AYQ_ONE_ISSUER=1 timeout 200 python tools/exp.py --tag nq1 --ops > gpurun_out/exp31_nq1.txt 2>&1
timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp31_base.txt 2>&1
grep -h "images/s" gpurun_out/exp31_*.txt

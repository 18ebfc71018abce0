#!/bin/bash
cd /root/repo
AYQ_PLAN_DUMP=1 timeout 300 python tools/exp.py --tag tuned --ops > gpurun_out/exp33_tuned.txt 2>&1
timeout 200 python tools/exp.py --tag tuned2 > gpurun_out/exp33_tuned2.txt 2>&1
AYQ_AUTOTUNE=0 timeout 200 python tools/exp.py --tag untuned > gpurun_out/exp33_untuned.txt 2>&1
grep -h "images/s" gpurun_out/exp33_*.txt; grep "^tune" gpurun_out/exp33_tuned.txt | grep -v "default$" | cut -c1-150

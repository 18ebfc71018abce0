#!/bin/bash
cd /root/repo
AYQ_PLAN_DUMP=1 timeout 200 python tools/exp.py --tag nsq4_nbq2 --ops > gpurun_out/exp27_a.txt 2>&1
AYQ_SLEEP_MIN_NSQ=0 AYQ_SLEEP_MIN_NBQ=0 timeout 200 python tools/exp.py --tag always --ops > gpurun_out/exp27_b.txt 2>&1
AYQ_SLEEP_MIN_NSQ=3 AYQ_SLEEP_MIN_NBQ=2 timeout 200 python tools/exp.py --tag nsq3_nbq2 --ops > gpurun_out/exp27_c.txt 2>&1
AYQ_SLEEP_MIN_NSQ=99 AYQ_SLEEP_MIN_NBQ=99 timeout 200 python tools/exp.py --tag never --ops > gpurun_out/exp27_d.txt 2>&1
AYQ_SLEEP_MIN_NSQ=4 AYQ_SLEEP_MIN_NBQ=3 timeout 200 python tools/exp.py --tag nsq4_nbq3 --ops > gpurun_out/exp27_e.txt 2>&1
AYQ_SLEEP_MIN_NSQ=8 AYQ_SLEEP_MIN_NBQ=2 timeout 200 python tools/exp.py --tag nsq8_nbq2 --ops > gpurun_out/exp27_f.txt 2>&1
grep -h "images/s" gpurun_out/exp27_*.txt

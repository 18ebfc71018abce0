#!/bin/bash
cd /root/repo
for i in 1 2; do
timeout 200 python tools/exp.py --steps 50 --tag tuned50_$i > gpurun_out/exp35_t$i.txt 2>&1
AYQ_AUTOTUNE=0 timeout 200 python tools/exp.py --steps 50 --tag untuned50_$i > gpurun_out/exp35_u$i.txt 2>&1
done
timeout 200 python tools/exp.py --steps 10 --tag tuned10 > gpurun_out/exp35_t10.txt 2>&1
AYQ_AUTOTUNE=0 timeout 200 python tools/exp.py --steps 10 --tag untuned10 > gpurun_out/exp35_u10.txt 2>&1
grep -h "images/s" gpurun_out/exp35_*.txt

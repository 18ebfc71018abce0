#!/bin/bash
cd /root/repo
for m in 1 2 4 8; do
AYQ_P1_CTA_MUL=$m timeout 200 python tools/exp.py --tag p1mul$m --ops > gpurun_out/exp26_$m.txt 2>&1; grep -h "images/s" gpurun_out/exp26_$m.txt; grep -o "absmax=[0-9.]* Conv_P1=[0-9.]*" gpurun_out/exp26_$m.txt
done

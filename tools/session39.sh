#!/bin/bash
cd /root/repo
for ns in 0 3000 6000 10000; do
AYQ_P1_STAGGER_NS=$ns timeout 200 python tools/exp.py --tag stagger$ns --ops > gpurun_out/exp39_$ns.txt 2>&1
grep -h "images/s" gpurun_out/exp39_$ns.txt; grep -o "Conv_P1=[0-9.]*" gpurun_out/exp39_$ns.txt
done

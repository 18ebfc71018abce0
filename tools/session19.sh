#!/bin/bash
cd /root/repo
AYQ_CARVEOUT=1 timeout 200 python tools/exp.py --tag carve > gpurun_out/exp19_carve.txt 2>&1
timeout 200 python tools/exp.py --tag base > gpurun_out/exp19_base.txt 2>&1
AYQ_CARVEOUT=1 AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/timeline_256_carve.txt 2>&1
tail -1 gpurun_out/timeline_256_carve.txt; grep -h "images/s" gpurun_out/exp19_*.txt

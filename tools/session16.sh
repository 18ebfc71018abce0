#!/bin/bash
cd /root/repo
AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_EPI_SKIP=1 timeout 200 python tools/exp.py --tag skip --ops > gpurun_out/exp16_skip.txt 2>&1
AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so timeout 200 python tools/exp.py --tag profbase --ops > gpurun_out/exp16_base.txt 2>&1
grep -h "images/s\|sum of" gpurun_out/exp16_*.txt

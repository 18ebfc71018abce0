#!/bin/bash
cd /root/repo
timeout 200 python tools/exp.py --tag namedbar --ops > gpurun_out/exp23.txt 2>&1; grep -h "images/s\|sum of" gpurun_out/exp23.txt
AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/timeline_256_s23.txt 2>&1; tail -1 gpurun_out/timeline_256_s23.txt
grep "entry-prev" gpurun_out/timeline_256_s23.txt | sed 's/.*first exit *\([-0-9.]*\),.*/\1/' | sort -n | awk '{a[NR]=$1} END {print "entry-prev first exit median", a[int(NR/2)], "min", a[1], "max", a[NR]}'

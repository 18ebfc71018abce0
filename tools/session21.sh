#!/bin/bash
cd /root/repo
for cfg in "1 2" "16 2" "0 33" "16 33" "1 33"; do
set -- $cfg
AYQ_EPI_SKIP=$1 AYQ_SKIP_TILES=$2 AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/timeline_x$1_$2.txt 2>&1
echo "mode=$1 tiles=$2"
grep "entry-prev" gpurun_out/timeline_x$1_$2.txt | sed 's/.*first exit *\([-0-9.]*\),.*/\1/' | sort -n | awk '{a[NR]=$1} END {print "entry-prev first exit median", a[int(NR/2)], "min", a[1], "max", a[NR]}'
tail -1 gpurun_out/timeline_x$1_$2.txt
done

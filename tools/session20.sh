#!/bin/bash
cd /root/repo
for k in 2 9 33; do
AYQ_SKIP_TILES=$k AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/timeline_256_skip$k.txt 2>&1
echo "k=$k"; grep "entry-prev" gpurun_out/timeline_256_skip$k.txt | awk '{print $NF}' | sort -n | awk '{a[NR]=$1} END {print "entry-prev last exit median", a[int(NR/2)]}'
grep "entry-prev" gpurun_out/timeline_256_skip$k.txt | sed 's/.*first exit *\([-0-9.]*\),.*/\1/' | sort -n | awk '{a[NR]=$1} END {print "entry-prev first exit median", a[int(NR/2)], "min", a[1], "max", a[NR]}'
tail -1 gpurun_out/timeline_256_skip$k.txt
done

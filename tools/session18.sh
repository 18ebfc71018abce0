#!/bin/bash
# kernel-to-kernel timeline (globaltimer stamps of the profiling build): where the fixed cost per launch goes
cd /root/repo
AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/timeline_256.txt 2>&1
AYQ_NO_PDL=1 AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/timeline_256_nopdl.txt 2>&1
AYQ_NO_PDL=1 timeout 200 python tools/exp.py --tag nopdl > gpurun_out/exp18_nopdl.txt 2>&1
timeout 200 python tools/exp.py --tag pdl > gpurun_out/exp18_pdl.txt 2>&1
tail -1 gpurun_out/timeline_256.txt; tail -1 gpurun_out/timeline_256_nopdl.txt; grep -h "images/s" gpurun_out/exp18_*.txt

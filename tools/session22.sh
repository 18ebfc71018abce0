#!/bin/bash
cd /root/repo
timeout 60 tools/ubench/pdlgap5 > gpurun_out/pdlgap5.txt 2>&1; cat gpurun_out/pdlgap5.txt | cut -c1-230
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_s22.txt 2>&1; tail -3 gpurun_out/pytest_s22.txt
timeout 200 python tools/exp.py --tag lutbulk --ops > gpurun_out/exp22.txt 2>&1; grep -h "images/s\|sum of" gpurun_out/exp22.txt
AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/timeline_256_s22.txt 2>&1; tail -1 gpurun_out/timeline_256_s22.txt

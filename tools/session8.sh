#!/bin/bash
cd /root/repo
timeout 100 python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/one_pass_plain8.txt 2>&1 && AYQ_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 62 -c 3 -o gpurun_out/conv_s8 -f python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/ncu_s8.log 2>&1
ls -la gpurun_out/conv_s8.ncu-rep; tail -3 gpurun_out/ncu_s8.log

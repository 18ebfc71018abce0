#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_s38.txt 2>&1; tail -3 gpurun_out/pytest_s38.txt
for i in 1 2; do timeout 200 python tools/exp.py --tag fma_gather_$i --ops > gpurun_out/exp38_$i.txt 2>&1; done
grep -h "images/s" gpurun_out/exp38_*.txt
grep -o "Conv_P1=[0-9.]* \|Conv_P2=[0-9.]* \|C2F_2_conv_0=[0-9.]* \|C2F_2_bottle_0=[0-9.]* \|C2F_2_conv_1=[0-9.]* \|C2F_15_conv_0=[0-9.]* \|x_result_5_down_1=[0-9.]* " gpurun_out/exp38_1.txt | tr '\n' ' '

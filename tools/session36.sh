#!/bin/bash
cd /root/repo
for i in 1 2 3; do timeout 200 python tools/exp.py --tag streamed_nq1_$i --ops > gpurun_out/exp36_$i.txt 2>&1; done
grep -h "images/s" gpurun_out/exp36_*.txt
grep -o "C2F_8_bottle_0=[0-9.]* \|C2F_8_bottle_1=[0-9.]* \|Conv_19=[0-9.]* \|C2F_21_bottle_0=[0-9.]* \|C2F_21_bottle_1=[0-9.]* \|x_up_0=[0-9.]* \|Conv_P5=[0-9.]* \|x_down_0=[0-9.]* \|C2F_8_conv_1=[0-9.]* \|SPPF_conv_1=[0-9.]* " gpurun_out/exp36_1.txt | tr '\n' ' '

#!/bin/bash
cd /root/repo
for i in 1 2; do timeout 200 python tools/exp.py --tag bhelper_$i --ops > gpurun_out/exp37_$i.txt 2>&1; done
AYQ_NO_B_HELPER=1 timeout 200 python tools/exp.py --tag nohelper --ops > gpurun_out/exp37_n.txt 2>&1
grep -h "images/s\|rror" gpurun_out/exp37_*.txt
for f in 1 n; do grep -o "C2F_8_bottle_0=[0-9.]* \|C2F_8_bottle_1=[0-9.]* \|Conv_19=[0-9.]* \|C2F_21_bottle_0=[0-9.]* \|C2F_21_bottle_1=[0-9.]* \|x_up_0=[0-9.]* \|Conv_P5=[0-9.]* \|x_down_0=[0-9.]* \|C2F_8_conv_1=[0-9.]* \|SPPF_conv_1=[0-9.]* " gpurun_out/exp37_$f.txt | tr '\n' ' '; echo; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_s37.txt 2>&1; tail -3 gpurun_out/pytest_s37.txt

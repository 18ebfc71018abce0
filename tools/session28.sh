#!/bin/bash
cd /root/repo
for kb in 150 190; do
AYQ_RESIDENT_KB=$kb AYQ_PLAN_DUMP=1 timeout 200 python tools/exp.py --tag res$kb --ops > gpurun_out/exp28_$kb.txt 2>&1
grep -h "images/s\|rror" gpurun_out/exp28_$kb.txt | head -5
grep "^plan" gpurun_out/exp28_$kb.txt | grep "resB=1" | awk '$8+0 >= 36 || $7 ~ /nkc=/ {print $2, $7, $8, $9, $12, $13, $14, $15, $16}' | sort -u | grep -i "nkc= *\(72\|144\|36\)" | head -20
grep -o "Conv_P5=[0-9.]* \|C2F_8_bottle_0=[0-9.]* \|C2F_8_bottle_1=[0-9.]* \|Conv_19=[0-9.]* \|C2F_21_bottle_0=[0-9.]* \|x_up_0=[0-9.]* \|x_down_0=[0-9.]* \|C2F_8_conv_1=[0-9.]* \|SPPF_conv_1=[0-9.]* " gpurun_out/exp28_$kb.txt | tr '\n' ' '; echo
done

#!/bin/bash
# pass-size lever (512 / 1024 images per pass), the 1-GPU batch sweep (BASELINE configs[3]) and a full capture of the non-conv kernels
cd /root/repo
timeout 200 python tools/exp.py --tag base256 > gpurun_out/exp17_256.txt 2>&1
timeout 200 python tools/exp.py --batch 512 --steps 6 --tag pass512 > gpurun_out/exp17_512.txt 2>&1
timeout 300 python tools/exp.py --batch 1024 --steps 4 --tag pass1024 > gpurun_out/exp17_1024.txt 2>&1
grep -h "images/s" gpurun_out/exp17_*.txt
timeout 400 python tools/batch_sweep.py --out gpurun_out/batch_sweep_r2.json > gpurun_out/batch_sweep_r2.txt 2>&1; tail -14 gpurun_out/batch_sweep_r2.txt
timeout 100 python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/one_pass_plain17.txt 2>&1 && AYQ_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:conv_p1|head_kernel|nms_kernel|sppf_pool|absmax' -s 5 -c 5 -o gpurun_out/misc_s17 -f python tools/one_pass.py --batch 256 --passes 2 > gpurun_out/ncu_s17.log 2>&1
ls -la gpurun_out/misc_s17.ncu-rep; tail -3 gpurun_out/ncu_s17.log

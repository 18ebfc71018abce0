#!/bin/bash
# GPU session (round 2): every step under its own timeout; nothing may hang the box
cd /root/repo
timeout 120 ./tools/ubench/ubench > gpurun_out/ubench_r2.txt 2>&1; echo "ubench rc=$?" >> gpurun_out/ubench_r2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "every_tensor and tma or all_golden and tma or host_entry or different_streams or nms_corner or full_size" > gpurun_out/pytest_s2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s2.txt
timeout 200 python tools/exp.py --tag base --ops > gpurun_out/exp_base.txt 2>&1
AYQ_NBUF_MUL=1 timeout 200 python tools/exp.py --tag nbuf1 --ops > gpurun_out/exp_nbuf1.txt 2>&1
AYQ_ROLE_HI=1 timeout 200 python tools/exp.py --tag role_hi --ops > gpurun_out/exp_rolehi.txt 2>&1
AYQ_ROLE_HI=1 AYQ_HALO_MIN_NP=1 timeout 200 python tools/exp.py --tag halo1_rolehi --ops > gpurun_out/exp_halo1_rolehi.txt 2>&1
AYQ_HALO_MIN_NP=1 timeout 200 python tools/exp.py --tag halo1 --ops > gpurun_out/exp_halo1.txt 2>&1
timeout 200 python tools/exp.py --tag dual --dual > gpurun_out/exp_dual.txt 2>&1
timeout 200 python tools/exp.py --tag b512 --batch 512 > gpurun_out/exp_b512.txt 2>&1
tail -n 3 gpurun_out/exp_*.txt gpurun_out/pytest_s2.txt; cat gpurun_out/ubench_r2.txt

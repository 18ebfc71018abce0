#!/bin/bash
# GPU session 1 (round 2): micro-benchmarks + cheap scheduling experiments
cd /root/repo
./tools/ubench/ubench > gpurun_out/ubench_r2.txt 2>&1
python tools/exp.py --tag base --ops > gpurun_out/exp_base.txt 2>&1
AYQ_ROLE_HI=1 python tools/exp.py --tag role_hi --ops > gpurun_out/exp_rolehi.txt 2>&1
python tools/exp.py --tag dual --dual > gpurun_out/exp_dual.txt 2>&1
AYQ_ROLE_HI=1 python tools/exp.py --tag dual_rolehi --dual > gpurun_out/exp_dual_rolehi.txt 2>&1
AYQ_HALO_MIN_NP=1 python tools/exp.py --tag halo1 --ops > gpurun_out/exp_halo1.txt 2>&1
AYQ_ROLE_HI=1 AYQ_HALO_MIN_NP=1 python tools/exp.py --tag halo1_rolehi --ops > gpurun_out/exp_halo1_rolehi.txt 2>&1
python tools/exp.py --tag b512 --batch 512 > gpurun_out/exp_b512.txt 2>&1
AYQ_ROLE_HI=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_rolehi.txt 2>&1
tail -3 gpurun_out/exp_*.txt gpurun_out/pytest_rolehi.txt; cat gpurun_out/ubench_r2.txt

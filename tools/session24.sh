#!/bin/bash
# final evidence of the round on the final build: GPU tests, bench lines (K = 8 / 6 / 4, 512-image passes), batch sweep, launch list, timeline
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final_r2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final_r2.txt; tail -n 3 gpurun_out/pytest_gpu_final_r2.txt
timeout 600 python bench.py --steps 20 --warmup 5 --sustain 3 --ops-json gpurun_out/ops_r2.json > gpurun_out/bench_1gpu_r2.json 2> gpurun_out/bench_1gpu_r2.err; tail -c 1500 gpurun_out/bench_1gpu_r2.json
timeout 300 python bench.py --steps 10 --warmup 3 --k 6 --no-cpu > gpurun_out/bench_k6_r2.json 2> gpurun_out/bench_k6_r2.err
timeout 300 python bench.py --steps 10 --warmup 3 --k 4 --no-cpu > gpurun_out/bench_k4_r2.json 2> gpurun_out/bench_k4_r2.err
timeout 300 python bench.py --steps 10 --warmup 3 --batch 512 --max-batch 512 --no-cpu > gpurun_out/bench_pass512_r2.json 2> gpurun_out/bench_pass512_r2.err
for f in k6 k4 pass512; do python -c "import json,sys; d=json.loads(open('gpurun_out/bench_${f}_r2.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['roofline']['frac'])"; done
timeout 400 python tools/batch_sweep.py --out gpurun_out/batch_sweep_r2.json > gpurun_out/batch_sweep_r2.txt 2>&1
timeout 400 python tools/batch_sweep.py --max-batch 512 --sizes 256,512,1024,2048,4096 --out gpurun_out/batch_sweep_pass512_r2.json > gpurun_out/batch_sweep_pass512_r2.txt 2>&1; tail -5 gpurun_out/batch_sweep_pass512_r2.txt
AYQ_LIB=alpha_yolo_quant_b200/libayq_prof.so AYQ_ROLE_PROF=1 timeout 200 python tools/one_pass.py --batch 256 --passes 3 --conv tma > gpurun_out/role_profile_r2.txt 2>&1; tail -1 gpurun_out/role_profile_r2.txt
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_plain_r2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 140 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_r2.log 2>&1
tail -2 gpurun_out/ncu_launches_r2.log

#!/bin/bash
# Run on the GPU box (gpurun): regenerates the evidence under gpurun_out/ that is then copied into profiles/.
#   1. plain bench run (must exit 0), 2. per-launch time list of the same command, 3. full capture of the dominant kernel,
#   4. role profile, 5. clocks during the bench.
set -x
R=${1:-r1}
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_plain_$R.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 140 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_$R.log 2>&1
python tools/one_pass.py --batch 256 --passes 2 --conv tma > gpurun_out/one_pass_plain_$R.log 2>&1 || exit 1
# DRAM traffic / pipe utilisation of every conv launch of one pass (few metrics = few replays)
AYQ_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size \
    --clock-control none -k regex:conv_tma -s 62 -c 62 --csv --page raw \
    --log-file gpurun_out/conv_tma_full_$R.csv python tools/one_pass.py --batch 256 --passes 2 --conv tma > gpurun_out/ncu_full_$R.log 2>&1
# full-set capture (with source) of three representative launches: a 1x1 conv, a halo 3x3 conv, a stride-2 conv
AYQ_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 67 -c 3 -o gpurun_out/conv_tma_set_full_$R \
    python tools/one_pass.py --batch 256 --passes 2 --conv tma > gpurun_out/ncu_set_full_$R.log 2>&1
AYQ_ROLE_PROF=1 python tools/one_pass.py --batch 256 --passes 2 --conv tma > gpurun_out/role_profile_$R.txt 2>&1
python bench.py --steps 20 --warmup 5 --ops-json gpurun_out/ops_$R.json > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 600 gpurun_out/bench_$R.json

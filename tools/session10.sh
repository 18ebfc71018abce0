#!/bin/bash
cd /root/repo
timeout 600 python bench.py --steps 20 --warmup 5 --sustain 3 --ops-json gpurun_out/ops_r2.json > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --k 6 --no-cpu > gpurun_out/bench_k6_r2.json 2> gpurun_out/bench_k6_r2.err
timeout 300 python bench.py --steps 20 --warmup 5 --k 4 --no-cpu > gpurun_out/bench_k4_r2.json 2> gpurun_out/bench_k4_r2.err
timeout 100 python tools/h2d_probe.py gpurun_out/h2d_probe_1gpu.json > gpurun_out/h2d_probe_1gpu.txt 2>&1
tail -c 3000 gpurun_out/bench_r2.json; tail -5 gpurun_out/bench_r2.err; tail -c 600 gpurun_out/bench_k6_r2.json; cat gpurun_out/h2d_probe_1gpu.txt

#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_s30.txt 2>&1; tail -3 gpurun_out/pytest_s30.txt
timeout 600 python bench.py --steps 20 --warmup 5 --sustain 3 --ops-json gpurun_out/ops_r2.json > gpurun_out/bench_1gpu_r2.json 2> gpurun_out/bench_1gpu_r2.err; tail -3 gpurun_out/bench_1gpu_r2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_1gpu_r2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'u8 resident', d['value_u8_resident'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
P

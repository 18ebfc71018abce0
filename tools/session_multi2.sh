#!/bin/bash
# 8-GPU box, final build of the round: multi-GPU parity test, bench at N = 2, 4, 8 (torchrun, one rank per GPU)
cd /root/repo
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi_r2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_r2.txt
for N in 2 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_${N}gpu_r2.json 2> gpurun_out/bench_${N}gpu_r2.err
done
tail -3 gpurun_out/pytest_multi_r2.txt
python - <<'P'
import json
for n in (2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/bench_{n}gpu_r2.json').read().strip().splitlines()[-1])
        print(n,'value',round(d['value']),'e2e u8',round(d['e2e']['value']),'f32',round(d['e2e_f32']['value']),'frac_h2d',d['e2e'].get('frac_of_h2d_probe'),'parity',d.get('multi_gpu_parity'))
    except Exception as ex: print(n,'fail',ex)
P

#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "host_entry or different_streams" > gpurun_out/pytest_s11.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s11.txt
for hp in 64 128 256; do
AYQ_HOST_PASS=$hp timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench11_hp$hp.json 2> gpurun_out/bench11_hp$hp.err
done
tail -3 gpurun_out/pytest_s11.txt
python - <<'P'
import json
for hp in (64,128,256):
    try:
        d=json.loads(open(f'gpurun_out/bench11_hp{hp}.json').read().strip().splitlines()[-1])
        print(hp,'value',round(d['value']),'e2e u8',round(d['e2e']['value']),'sync',round(d['e2e']['sync_value']),'f32',round(d['e2e_f32']['value']),'sync',round(d['e2e_f32']['sync_value']))
    except Exception as ex: print(hp,'fail',ex)
P

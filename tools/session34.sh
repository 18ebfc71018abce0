#!/bin/bash
cd /root/repo
for i in 1 2; do
AYQ_PLAN_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/b34_tuned$i.json 2> gpurun_out/b34_tuned$i.err
AYQ_AUTOTUNE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/b34_untuned$i.json 2> gpurun_out/b34_untuned$i.err
done
for f in tuned1 untuned1 tuned2 untuned2; do python -c "import json; d=json.loads(open('gpurun_out/b34_$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'])"; done
grep "^tune" gpurun_out/b34_tuned1.err | grep -v "default$" | cut -c1-120

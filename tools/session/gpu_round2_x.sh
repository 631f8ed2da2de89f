#!/bin/bash
# quickest solve-kernel check: the C2 bench line with its parity gate only
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-offline --no-c5 > gpurun_out/x_bench.log 2>&1; echo "bench rc=$?"
grep -a '^{' gpurun_out/x_bench.log | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']
print('value',l['value'],'ms_per_step',l['ms_per_step'],'kernel',r['kernel'],'ms_per_launch',r['ms_per_launch'],'frac',r['frac'],'parity',l['parity']['max_rel'])"
tail -2 gpurun_out/x_bench.log | cut -c1-160

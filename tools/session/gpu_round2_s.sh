#!/bin/bash
# solve_kernel_v2 experiment cycle: kernel tests, C2 bench line, phase counters (developer build last)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "solve or reproducible or estimate" 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-offline --no-c5 > gpurun_out/s_bench.log 2>&1; echo "bench rc=$?"
grep -a '^{' gpurun_out/s_bench.log | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']
print('value',l['value'],'ms_per_step',l['ms_per_step'],'kernel',r['kernel'],'ms_per_launch',r['ms_per_launch'],'frac',r['frac'],'exec_frac',r.get('executed_frac'),'parity',l.get('parity'))"
tail -2 gpurun_out/s_bench.log | cut -c1-300
LRBMS_DEVTOOLS=1 python -m pylrbms_b200.build --force > /dev/null && timeout 300 python tools/solve_timing.py --solver window > gpurun_out/s_timing_v2.txt 2>&1; echo "timing rc=$?"
tail -22 gpurun_out/s_timing_v2.txt

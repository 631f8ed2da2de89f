#!/bin/bash
# full regression: GPU test suite, smoke(), default bench line (kept as r02c_bench_1gpu.json), reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/v_bench.log 2>&1; echo "bench rc=$?"
grep -a '^{' gpurun_out/v_bench.log | tail -1 > gpurun_out/r02c_bench_1gpu.json
python - <<PY
import json
l=json.load(open('gpurun_out/r02c_bench_1gpu.json'))
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'], l['parity']['max_rel'], l['clocks'], (l.get('c5_sweep') or {}).get('value'), l['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600

#!/bin/bash
# validate the sparse-solve gate of bench.py on c3 / c4
mkdir -p gpurun_out
timeout 600 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-offline > gpurun_out/p_c3.log 2>&1; echo "c3 rc=$?"
tail -c 1500 gpurun_out/p_c3.log
timeout 900 python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-offline > gpurun_out/p_c4.log 2>&1; echo "c4 rc=$?"
tail -c 1500 gpurun_out/p_c4.log

set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "reproducible or barrier_schedule" > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02l_pytest.log
python bench.py --steps 5 --warmup 3 --no-offline --no-cpu-baseline --no-c5 --no-parity > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02l_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02l_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['kernel'], r['ms_per_launch'], r['frac'], r['executed_frac'])
"

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "band or solve_matches_dense" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02m_pytest.log
python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-offline > gpurun_out/r02m_bench_c3.json 2> gpurun_out/r02m_bench_c3.err; echo "c3 rc=$?"
python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-offline > gpurun_out/r02m_bench_c4.json 2> gpurun_out/r02m_bench_c4.err; echo "c4 rc=$?"
python -c "
import json
for f in ('c3','c4'):
    d=json.load(open('gpurun_out/r02m_bench_%s.json'%f)); r=d['roofline']; print(f, d['value'], r['ms_per_launch'], r['executed_frac'], r['frac'])
"

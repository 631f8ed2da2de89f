set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_adaptive_enrichment.py -m gpu -q -x -k "band or adaptive or batched or pcg or solve_matches" -s > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"
grep -i "enrichment parity\|passed\|failed\|error" gpurun_out/r02h_pytest.log | tail -8
python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-offline > gpurun_out/r02h_bench_c3.json 2> gpurun_out/r02h_bench_c3.err; echo "c3 rc=$?"
python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-offline > gpurun_out/r02h_bench_c4.json 2> gpurun_out/r02h_bench_c4.err; echo "c4 rc=$?"
python -c "
import json
for f in ('c3','c4'):
    d=json.load(open('gpurun_out/r02h_bench_%s.json'%f)); r=d['roofline']; print(f, d['value'], r['ms_per_launch'], r['executed_frac'], r['frac'])
"
python tools/enrichment_timing.py > gpurun_out/r02h_enrichment_timing.txt 2>&1; echo "enrich rc=$?"; tail -6 gpurun_out/r02h_enrichment_timing.txt
python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "c4" > gpurun_out/r02h_pytest_c4.log 2>&1; echo "pytest c4 rc=$?"; tail -3 gpurun_out/r02h_pytest_c4.log

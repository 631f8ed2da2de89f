set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py -m gpu -q -k "c1 or c2 or 4x4" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -15 gpurun_out/r02c_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/r02c_bench.err
python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_c3.json 2> gpurun_out/r02c_bench_c3.err; echo "bench c3 rc=$?"
tail -c 2000 gpurun_out/r02c_bench_c3.err
python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_c4.json 2> gpurun_out/r02c_bench_c4.err; echo "bench c4 rc=$?"
tail -c 2000 gpurun_out/r02c_bench_c4.err

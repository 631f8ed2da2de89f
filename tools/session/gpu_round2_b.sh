set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -40 gpurun_out/r02b_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/r02b_bench.err

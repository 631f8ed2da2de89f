set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -25 gpurun_out/r02n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02n_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/r02n_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err; echo "ref rc=$?"

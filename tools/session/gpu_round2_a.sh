set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=15 -x > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -30 gpurun_out/r02a_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench.err
python tools/profile_reduce.py > gpurun_out/r02a_profile_reduce.txt 2>&1; echo "profile rc=$?"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_c1.py > gpurun_out/r02a_memcheck_c1.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r02a_memcheck_c1.log
CELLS=8 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_c1.py > gpurun_out/r02a_racecheck_c1.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/r02a_racecheck_c1.log

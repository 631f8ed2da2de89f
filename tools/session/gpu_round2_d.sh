set -x
mkdir -p gpurun_out
python tools/profile_reduce.py > gpurun_out/r02d_profile_reduce.txt 2>&1; echo "profile rc=$?"
grep "reduce() call\|discretize" gpurun_out/r02d_profile_reduce.txt
SUBDOMAINS=16 python tools/profile_reduce.py > gpurun_out/r02d_profile_reduce_c3.txt 2>&1; echo "profile rc=$?"
grep "reduce() call\|discretize" gpurun_out/r02d_profile_reduce_c3.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_configs.py -m gpu -q -x -k "not 16x16" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02d_pytest.log

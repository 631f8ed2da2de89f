set -x
mkdir -p gpurun_out
python tools/profile_reduce.py > gpurun_out/r02e_profile_reduce.txt 2>&1; echo "profile rc=$?"
grep "reduce() call\|discretize" gpurun_out/r02e_profile_reduce.txt
SUBDOMAINS=16 python tools/profile_reduce.py > gpurun_out/r02e_profile_reduce_c3.txt 2>&1; echo "profile rc=$?"
grep "reduce() call\|discretize" gpurun_out/r02e_profile_reduce_c3.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q -x -k "reduce or incremental or project" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02e_pytest.log

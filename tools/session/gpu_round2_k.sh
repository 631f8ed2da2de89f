set -x
mkdir -p gpurun_out
LRBMS_DEVTOOLS=1 python -m pylrbms_b200.build --force > /dev/null && python tools/solve_timing.py --solver panel > gpurun_out/r02k_timing_v3.txt 2>&1; echo "rc=$?"
cat gpurun_out/r02k_timing_v3.txt | tail -24

set -x
mkdir -p gpurun_out
LRBMS_DEVTOOLS=1 python -m pylrbms_b200.build --force > /dev/null && python tools/solve_timing.py --solver window > gpurun_out/r02o_timing_v2.txt 2>&1; echo "rc=$?"
cat gpurun_out/r02o_timing_v2.txt | tail -24

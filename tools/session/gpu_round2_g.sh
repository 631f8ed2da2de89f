set -x
mkdir -p gpurun_out
python tools/diag_determinism.py 2>&1 | tail -8
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reproducible or reduce_matches" 2>&1 | tail -3

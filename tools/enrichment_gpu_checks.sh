#!/bin/bash
# cooperative CG kernel + peer-push unit tests + enrichment step timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_adaptive_enrichment.py tests/test_gpu_kernels.py -m gpu -x -q -k "pcg or enrichment or peer_push or correction" 2>&1 | tail -8
timeout 400 python tools/enrichment_timing.py > gpurun_out/t_enrichment_timing.txt 2>&1; echo "timing rc=$?"
grep -a -v "^estimated error" gpurun_out/t_enrichment_timing.txt | tail -12

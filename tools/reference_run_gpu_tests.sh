#!/bin/bash
# new reference-run tests on the GPU
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_adaptive_enrichment.py -m gpu -x -q -s 2>&1 | grep -v "estimated error" | tail -12

#!/bin/bash
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_adaptive_enrichment.py -m gpu -x -q 2>&1 | tail -3

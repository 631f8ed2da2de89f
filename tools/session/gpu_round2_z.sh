#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/profile_enrichment.py > gpurun_out/z_profile_enrichment.txt 2>&1; echo "rc=$?"; grep -a "^step" gpurun_out/z_profile_enrichment.txt

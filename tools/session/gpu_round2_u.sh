#!/bin/bash
mkdir -p gpurun_out
nproc; uptime
for i in 1 2; do
  timeout 400 python tools/enrichment_timing.py > gpurun_out/u_enrichment_timing_$i.txt 2>&1; echo "timing rc=$?"
  grep -a -v "^estimated error" gpurun_out/u_enrichment_timing_$i.txt | tail -8
  uptime
done

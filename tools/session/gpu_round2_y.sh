#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pcg_timing.py > gpurun_out/y_pcg_timing.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/y_pcg_timing.txt

#!/bin/bash
# End-of-round measurement set (run on the GPU box through gpurun; outputs under gpurun_out/, copied to profiles/ by hand).
# Bench numbers are taken without a profiler; the ncu passes repeat the same commands afterwards.
set -u
TAG=${1:-r01b}
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err
python bench.py --steps 2 --warmup 3 --subdomains 16 --n-mu 296 --no-cpu-baseline > $O/${TAG}_bench_c3_16x16.json 2> $O/${TAG}_bench_c3.err
python bench.py --subdomains 4 --synthetic3d 16,16,12 --basis 40 --no-cpu-baseline > $O/${TAG}_offline_c4shape_4x4x4.json 2> $O/${TAG}_c4.err
python tools/incremental_timing.py > $O/${TAG}_incremental_timing.txt 2>&1
# launch list of the bench command (cold-cache, serialised: only the shares are comparable with the bench)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
# full captures: one run of the projection plan (single stream so that the launches do not overlap) ...
REPS=1 LRBMS_SINGLE_STREAM=1 ncu --set full --clock-control none --import-source on \
    -k regex:"project_kernel|spmm_kernel|gram_kernel" -s 72 -c 18 -f -o $O/${TAG}_project python tools/offline_timing.py > $O/${TAG}_ncu_project.log 2>&1
# ... and the two online kernels
ncu --set full --clock-control none --import-source on -k regex:"solve_kernel_v2" -s 3 -c 1 -f -o $O/${TAG}_solve \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-offline > $O/${TAG}_ncu_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"estimate_kernel" -s 3 -c 1 -f -o $O/${TAG}_estimate \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-offline > $O/${TAG}_ncu_estimate.log 2>&1
tail -c 600 $O/${TAG}_bench_1gpu.json; echo; tail -c 400 $O/${TAG}_offline_c4shape_4x4x4.json; echo; cat $O/${TAG}_incremental_timing.txt | tail -9

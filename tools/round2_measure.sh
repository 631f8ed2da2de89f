#!/bin/bash
# Round-2 measurement set (one gpurun call, 1 GPU).  Bench numbers are taken without a profiler; every ncu pass repeats a
# command that has just exited 0 without ncu.  Outputs under gpurun_out/, summaries copied to profiles/ by tools/ncu_summary.py.
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err; echo "bench rc=$?"
python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "c3 rc=$?"
python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "c4 rc=$?"
python bench.py --subdomains 4 --synthetic3d 16,16,12 --basis 40 --offline-only --no-cpu-baseline > $O/${TAG}_offline_c4shape_4x4x4.json 2> $O/${TAG}_c4shape.err; echo "c4shape rc=$?"
# launch list of the bench command (cold-cache, serialised: only the shares are comparable with the bench)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-c5"
$CMD > $O/${TAG}_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_steps2.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1
echo "launches rc=$?"
# full captures: one run of the projection plan on a single stream ...
REPS=1 SINGLE_STREAM=1 python tools/offline_timing.py > $O/${TAG}_plain2.log 2>&1 && \
REPS=1 SINGLE_STREAM=1 ncu --set full --clock-control none --import-source on \
    -k regex:"project_kernel|spmm_kernel|gram_kernel" -s 72 -c 18 -f -o $O/${TAG}_project python tools/offline_timing.py > $O/${TAG}_ncu_project.log 2>&1
echo "project rc=$?"
# ... the online kernels at C2 ...
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-offline --no-parity --no-c5"
$CMD2 > $O/${TAG}_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"solve_kernel_v2|estimate_kernel" -s 6 -c 2 -f -o $O/${TAG}_online $CMD2 > $O/${TAG}_ncu_online.log 2>&1
echo "online rc=$?"
# ... and the band solver at the 8x8x8, N = 40 size (block column 160 of 320: full-length updates; two update launches per
# column: the off-diagonal targets on the main stream, the diagonal target on the side stream)
CMD3="python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --no-offline"
$CMD3 > $O/${TAG}_plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"band_update_kernel" -s 321 -c 2 -f -o $O/${TAG}_band_update $CMD3 > $O/${TAG}_ncu_band.log 2>&1
echo "band rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"band_potrf_kernel|band_trsm_kernel|band_substitute_kernel" -s 320 -c 2 -f -o $O/${TAG}_band_rest $CMD3 > $O/${TAG}_ncu_band2.log 2>&1
echo "band2 rc=$?"
ls -la $O | tail -30

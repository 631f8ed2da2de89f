set -x
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02f_bench_${N}gpu.json 2> gpurun_out/r02f_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02f_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py > gpurun_out/r02f_check_sharded_${N}.txt 2>&1; echo "check_sharded rc=$?"; tail -3 gpurun_out/r02f_check_sharded_${N}.txt

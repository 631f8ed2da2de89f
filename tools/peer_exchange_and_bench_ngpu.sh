#!/bin/bash
# N GPUs: peer-memory exchange check/timing in every mode, then the full bench line
N=${1:-8}
mkdir -p gpurun_out
for mode in auto unicast multicast; do
  LRBMS_PEER_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/check_peer_exchange.py > gpurun_out/r_peer_${mode}_$N.log 2>&1; echo "peer check ($mode) rc=$?"
  grep -a '^{' gpurun_out/r_peer_${mode}_$N.log | tail -1
  grep -a -i "error\|Traceback" gpurun_out/r_peer_${mode}_$N.log | head -5
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r_bench_$N.log 2>&1; echo "bench rc=$?"
grep -a '^{' gpurun_out/r_bench_$N.log | tail -1 > gpurun_out/r02c_bench_${N}gpu.json
python - <<PY
import json
l=json.load(open('gpurun_out/r02c_bench_${N}gpu.json')); o=l.get('offline') or {}
print(json.dumps({k:o.get(k) for k in ('sharded','sharded_c3')}, indent=1))
print({k:l.get(k) for k in ('value','ms_per_step','n_gpus')}, l.get('e2e'), l.get('c5_sweep'))
PY
tail -3 gpurun_out/r_bench_$N.log | cut -c1-300

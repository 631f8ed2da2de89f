#!/bin/bash
# peer-memory exchange: bitwise check against NCCL + timing, then the sharded offline bench lines (run with --gpus N)
N=${1:-2}
mkdir -p gpurun_out
for mode in auto unicast; do
  LRBMS_PEER_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/check_peer_exchange.py > gpurun_out/q_peer_${mode}_$N.log 2>&1; echo "peer check ($mode) rc=$?"
  grep -a '^{' gpurun_out/q_peer_${mode}_$N.log | tail -1
  grep -a -i "error\|Traceback" gpurun_out/q_peer_${mode}_$N.log | head -5
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 5 --warmup 3 --offline-only --no-cpu-baseline > gpurun_out/q_bench_offline_$N.log 2>&1; echo "bench rc=$?"
grep -a '^{' gpurun_out/q_bench_offline_$N.log | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); o=l.get('offline') or {}
print(json.dumps({k:o.get(k) for k in ('sharded','sharded_c3')}, indent=1))"
tail -5 gpurun_out/q_bench_offline_$N.log | cut -c1-400

#!/bin/bash
# 8-GPU box: bench at N = 8 and N = 4 with the host-buffer e2e leg and the sharded C5 alignment
cd "$(dirname "$0")/.."
O=gpurun_out
for n in 8; do
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 > $O/r5q_bench_n$n.json 2> $O/r5q_bench_n$n.err
echo "n$n rc=$? wall=${SECONDS}s" >> $O/r5q_bench_n$n.err
tail -1 $O/r5q_bench_n$n.err; cut -c1-300 $O/r5q_bench_n$n.json
done

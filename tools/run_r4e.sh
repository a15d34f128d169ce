#!/bin/bash
# 2-GPU box: the multi-GPU tests (NCCL), N=2 bench with the host-buffer e2e leg after the grouped statistics of khg_estep
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_multirank_gloo.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > $O/r5f_gputests.txt
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r5f_bench_n2.json 2> $O/r5f_bench_n2.err
echo "n2 rc=$? wall=${SECONDS}s" >> $O/r5f_bench_n2.err
tail -3 $O/r5f_gputests.txt; tail -3 $O/r5f_bench_n2.err; cut -c1-1200 $O/r5f_bench_n2.json

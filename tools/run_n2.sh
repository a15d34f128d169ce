#!/bin/bash
# 2-GPU box: GPU tests (incl. the NCCL test), K3 check after the restore, N=2 bench with the host-buffer e2e leg, reference arm at N=2
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r3b_gputests.txt
timeout 120 python tools/bench_stats.py c4 > $O/r3b_bench_stats.txt 2>&1
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r3b_bench_n2.json 2> $O/r3b_bench_n2.err
echo "n2 rc=$? wall=${SECONDS}s" >> $O/r3b_bench_n2.err
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $O/r3b_bench_ref_n2.json 2> $O/r3b_bench_ref_n2.err
echo "ref rc=$? wall=${SECONDS}s" >> $O/r3b_bench_ref_n2.err
tail -5 $O/r3b_gputests.txt; cut -c1-250 $O/r3b_bench_stats.txt; tail -3 $O/r3b_bench_n2.err; cut -c1-1500 $O/r3b_bench_n2.json; tail -2 $O/r3b_bench_ref_n2.err; cut -c1-600 $O/r3b_bench_ref_n2.json

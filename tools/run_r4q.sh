#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4q_*
for r in 1 2 3; do
  KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 4 --check 0 > $O/r4q_run$r.txt 2>&1
  grep -o '"value_device_feats": [0-9.]*\|"e2e_host_feats": [0-9.]*' $O/r4q_run$r.txt | paste -s -d' '
  grep "khg_align_batch:" $O/r4q_run$r.txt | grep -o "host prep.*" | head -8
done

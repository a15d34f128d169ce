#!/bin/bash
# W-aligned on a C4-sized model with ragged pdf sizes (4..44 Gaussians): tensor-core kernel + fp32 list kernel for the large pdfs
# against the fp32 kernel alone (what such a model ran before), and against the previous build
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r5k_*
for v in "in-tree" "simt" "prev"; do
  echo "## $v" >> $O/r5k_ragged.txt
  if [ $v = simt ]; then KHG_STATS_KERNEL=simt KHG_BENCH_SIZES=4,44 timeout 200 python tools/bench_stats.py c4 >> $O/r5k_ragged.txt 2>&1
  elif [ $v = prev ]; then KHG_B200_LIB=tools/ab/stk_prev.so KHG_BENCH_SIZES=4,44 timeout 200 python tools/bench_stats.py c4 >> $O/r5k_ragged.txt 2>&1
  else KHG_BENCH_SIZES=4,44 timeout 200 python tools/bench_stats.py c4 >> $O/r5k_ragged.txt 2>&1; fi
done
cut -c1-420 $O/r5k_ragged.txt

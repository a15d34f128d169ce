#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2w_*.txt
timeout 600 python -m pytest tests/test_gpu_stats_tc.py -m gpu -q -x 2>&1 | tail -3 > $O/r2w_tests.txt
KHG_STATS_TC_CTAS_PER_SM=1 KHG_B200_LIB=tools/ab/stk_timing.so timeout 120 python tools/bench_stats.py c4 2>&1 | grep -v "^{" | head -1 >> $O/r2w_timing.txt
KHG_B200_LIB=tools/ab/stk_timing.so timeout 120 python tools/bench_stats.py c4 2>&1 | grep -v "^{" | head -1 >> $O/r2w_timing.txt
for r in 1 2; do
echo "first 4 (in-tree)" >> $O/r2w_bench_stats.txt; timeout 120 python tools/bench_stats.py c4 >> $O/r2w_bench_stats.txt 2>&1
for k in 2 6 16; do echo "first $k" >> $O/r2w_bench_stats.txt; KHG_B200_LIB=tools/ab/stk_b$k.so timeout 120 python tools/bench_stats.py c4 >> $O/r2w_bench_stats.txt 2>&1; done; done
cat $O/r2w_tests.txt $O/r2w_timing.txt; cut -c1-150 $O/r2w_bench_stats.txt

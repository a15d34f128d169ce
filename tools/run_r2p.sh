#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2p_*.txt
for c in c3 c4 c5; do timeout 120 python tools/stats_tc_check.py $c 400000 >> $O/r2p_check.txt 2>&1 || echo "FAILED/timeout $c rc=$?" >> $O/r2p_check.txt; done
for r in 1 2; do
for c in c4 c3; do echo "packed $c" >> $O/r2p_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r2p_bench_stats.txt 2>&1; 
echo "prev $c" >> $O/r2p_bench_stats.txt; KHG_B200_LIB=tools/ab/stk_prev.so timeout 120 python tools/bench_stats.py $c >> $O/r2p_bench_stats.txt 2>&1; done; done
tail -12 $O/r2p_check.txt | cut -c1-330; cut -c1-200 $O/r2p_bench_stats.txt

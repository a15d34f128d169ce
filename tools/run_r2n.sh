#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2n_*.txt
for c in c3 c4 c2 c5; do KHG_STATS_TC_DEBUG=1 timeout 120 python tools/stats_tc_check.py $c 400000 >> $O/r2n_check.txt 2>&1 || echo "FAILED/timeout $c rc=$?" >> $O/r2n_check.txt; done
STK_DETAIL=1 timeout 120 python tools/stats_tc_check.py c4 2000000 2>&1 | tail -22 > $O/r2n_detail.txt
for c in c4 c3 c2 c5; do echo "in-tree (prefetch 1, 3 CTAs) $c" >> $O/r2n_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r2n_bench_stats.txt 2>&1; done
for v in stk_p0_c3 stk_p0_c4 stk_p1_c2; do echo "$v" >> $O/r2n_bench_stats.txt; KHG_B200_LIB=tools/ab/$v.so timeout 120 python tools/bench_stats.py c4 >> $O/r2n_bench_stats.txt 2>&1; done
echo "simt" >> $O/r2n_bench_stats.txt; KHG_STATS_KERNEL=simt timeout 120 python tools/bench_stats.py c4 >> $O/r2n_bench_stats.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_tc_kernel -s 2 -c 1 -o $O/r2n_stats_tc python tools/bench_stats.py c4 2000000 > $O/r2n_ncu.log 2>&1
grep -v "^stats_tc_kernel" $O/r2n_check.txt | tail -12 | cut -c1-400; grep "^stats_tc_kernel" $O/r2n_check.txt | sort | uniq -c; cat $O/r2n_detail.txt | cut -c1-300; cut -c1-200 $O/r2n_bench_stats.txt

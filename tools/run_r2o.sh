#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2o_*.txt
for c in c3 c4 c2 c5; do timeout 120 python tools/stats_tc_check.py $c 400000 >> $O/r2o_check.txt 2>&1 || echo "FAILED/timeout $c rc=$?" >> $O/r2o_check.txt; done
for c in c4 c3 c2 c5; do echo "tc $c" >> $O/r2o_bench_stats.txt; timeout 120 python tools/bench_stats.py $c >> $O/r2o_bench_stats.txt 2>&1; done
echo "simt" >> $O/r2o_bench_stats.txt; KHG_STATS_KERNEL=simt timeout 120 python tools/bench_stats.py c4 >> $O/r2o_bench_stats.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_tc_kernel -s 2 -c 1 -o $O/r2o_stats_tc python tools/bench_stats.py c4 2000000 > $O/r2o_ncu.log 2>&1
tail -12 $O/r2o_check.txt | cut -c1-330; cut -c1-200 $O/r2o_bench_stats.txt

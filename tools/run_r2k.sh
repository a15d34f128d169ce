#!/bin/bash
# first run of the tensor-core statistics kernel (K3t): check against the fp32 kernel, then rates
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2k_*.txt
for c in c3 c4 c2 c5; do timeout 120 python tools/stats_tc_check.py $c 400000 >> $O/r2k_check.txt 2>&1 || echo "FAILED/timeout $c rc=$?" >> $O/r2k_check.txt; done
for k in simt tc; do for c in c4 c3 c2 c5; do echo "kernel $k" >> $O/r2k_bench_stats.txt; KHG_STATS_KERNEL=$k timeout 120 python tools/bench_stats.py $c >> $O/r2k_bench_stats.txt 2>&1; done; done
for n in 1 3; do echo "tc ctas_per_sm $n" >> $O/r2k_bench_stats.txt; KHG_STATS_TC_CTAS_PER_SM=$n timeout 120 python tools/bench_stats.py c4 >> $O/r2k_bench_stats.txt 2>&1; done
tail -12 $O/r2k_check.txt | cut -c1-400; cut -c1-200 $O/r2k_bench_stats.txt

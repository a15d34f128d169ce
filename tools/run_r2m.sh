#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2m_*.txt
for c in c3 c4 c2 c5; do KHG_STATS_TC_DEBUG=1 timeout 120 python tools/stats_tc_check.py $c 400000 >> $O/r2m_check.txt 2>&1 || echo "FAILED/timeout $c rc=$?" >> $O/r2m_check.txt; done
for k in tc; do for c in c4 c3 c2 c5; do echo "kernel $k" >> $O/r2m_bench_stats.txt; KHG_STATS_KERNEL=$k timeout 120 python tools/bench_stats.py $c >> $O/r2m_bench_stats.txt 2>&1; done; done
for n in 1 3; do echo "tc ctas_per_sm $n" >> $O/r2m_bench_stats.txt; KHG_STATS_TC_CTAS_PER_SM=$n timeout 120 python tools/bench_stats.py c4 >> $O/r2m_bench_stats.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_tc_kernel -s 2 -c 1 -o $O/r2m_stats_tc python tools/bench_stats.py c4 2000000 > $O/r2m_ncu.log 2>&1
grep -v "^stats_tc_kernel" $O/r2m_check.txt | tail -12 | cut -c1-400; grep "^stats_tc_kernel" $O/r2m_check.txt | sort | uniq -c; cut -c1-200 $O/r2m_bench_stats.txt

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stats_tc_kernel -s 2 -c 1 -o $O/r2l_stats_tc python tools/bench_stats.py c4 2000000 > $O/r2l_ncu.log 2>&1
tail -3 $O/r2l_ncu.log

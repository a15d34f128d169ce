#!/bin/bash
# round-2 experiment batch b: what bounds K1 (both forms) at ~51.5 M frames/s?
cd "$(dirname "$0")/.."
O=gpurun_out
tools/write_pattern.bin > $O/r2b_write_pattern.txt 2>&1
echo "## Gaussian-stationary kernel (experiments build): 0 normal, 1 no LSE, 2 A' tile re-read (hot L2), 10 stores into an L2-resident window, 11 no stores" > $O/r2b_gs_modes.txt
KHG_B200_LIB=tools/ab/exp.so timeout 200 python tools/k1_modes.py 3 0,1,2,10,11,0 >> $O/r2b_gs_modes.txt 2>&1
echo "## sub-block size KHG_GS_CHUNK" >> $O/r2b_gs_modes.txt
for c in 37888 75776 151552 303104; do echo "chunk $c" >> $O/r2b_gs_modes.txt; KHG_GS_CHUNK=$c timeout 100 python tools/k1_modes.py 3 0 >> $O/r2b_gs_modes.txt 2>&1; done
echo "## ring depth KHG_GS_STAGES" >> $O/r2b_gs_modes.txt
for s in 3 5; do echo "stages $s" >> $O/r2b_gs_modes.txt; KHG_GS_STAGES=$s timeout 100 python tools/k1_modes.py 3 0 >> $O/r2b_gs_modes.txt 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loglikes_gs_kernel -c 1 -o $O/r2b_gs python tools/prof_dense.py 75776 3 > $O/r2b_ncu.log 2>&1
cat $O/r2b_write_pattern.txt $O/r2b_gs_modes.txt; tail -3 $O/r2b_ncu.log

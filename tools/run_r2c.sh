#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 100 python tools/gs_check.py > $O/r2c_gs_check.txt 2>&1; echo "rc=$?" >> $O/r2c_gs_check.txt
echo "## K1g, rotation of the epilogue roles over frame tiles (KHG_GS_ROTATE), alternating; then KHG_GS=0 (frame-stationary kernel)" > $O/r2c_gs_rotate.txt
for r in 0 1 0 1; do echo "rotate $r" >> $O/r2c_gs_rotate.txt; KHG_GS_ROTATE=$r timeout 100 python tools/k1_modes.py 3 0 >> $O/r2c_gs_rotate.txt 2>&1; done
KHG_GS=0 timeout 100 python tools/k1_modes.py 3 0 >> $O/r2c_gs_rotate.txt 2>&1
echo "## experiments build: 0, 1 (no LSE), 10 (L2-window stores)" >> $O/r2c_gs_rotate.txt
KHG_B200_LIB=tools/ab/exp.so timeout 100 python tools/k1_modes.py 3 0,1,10 >> $O/r2c_gs_rotate.txt 2>&1
echo "## C5 / C3 / C2 (K1_CONFIG), GS then frame-stationary" >> $O/r2c_gs_rotate.txt
for c in c5 c3 c2; do echo "$c" >> $O/r2c_gs_rotate.txt; K1_CONFIG=$c timeout 100 python tools/k1_modes.py 3 0 >> $O/r2c_gs_rotate.txt 2>&1; K1_CONFIG=$c KHG_GS=0 timeout 100 python tools/k1_modes.py 3 0 >> $O/r2c_gs_rotate.txt 2>&1; done
cat $O/r2c_gs_check.txt $O/r2c_gs_rotate.txt

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
echo "## K1 (frame-stationary, f16 split), experiments build: 0 normal, 12 tile-major addressing inside the same block, 10 L2-window stores; twice" > $O/r2t_k1_tile_major.txt
for r in 1 2; do KHG_B200_LIB=tools/ab/exp.so timeout 300 python tools/k1_modes.py 3 0,12,10 >> $O/r2t_k1_tile_major.txt 2>&1; done
echo "## c5" >> $O/r2t_k1_tile_major.txt
KHG_B200_LIB=tools/ab/exp.so K1_CONFIG=c5 timeout 300 python tools/k1_modes.py 3 0,12 >> $O/r2t_k1_tile_major.txt 2>&1
echo "## c3" >> $O/r2t_k1_tile_major.txt
KHG_B200_LIB=tools/ab/exp.so K1_CONFIG=c3 timeout 300 python tools/k1_modes.py 3 0,12 >> $O/r2t_k1_tile_major.txt 2>&1
cat $O/r2t_k1_tile_major.txt

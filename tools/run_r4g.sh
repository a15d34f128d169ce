#!/bin/bash
# aligner search kernel with the graph in shared memory: threads per utterance x frames per tile at C5; tests first
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4g_*
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_em_recipe.py -m gpu -q -x 2>&1 | tail -4 > $O/r4g_tests.txt
cat $O/r4g_tests.txt
for cfg in "128 32 1" "128 8 1" "128 8 0" "64 8 1" "64 16 1" "32 8 1" "256 8 1"; do
  set -- $cfg
  echo "## NT=$1 FC=$2 GRAPH_SMEM=$3" >> $O/r4g_align_nt.txt
  KHG_ALIGN_NT=$1 KHG_ALIGN_FORCE_FC=$2 KHG_ALIGN_GRAPH_SMEM=$3 KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 4 --check 4 2>&1 | grep -v "^khg_align_batch host" | cut -c1-420 | tail -4 >> $O/r4g_align_nt.txt
done
grep -o "## NT.*\|smem [0-9]*\|search [0-9.]* ms\|\"value_device_feats\": [0-9.]*" $O/r4g_align_nt.txt | paste -s -d' ' | sed 's/## /\n/g'

#!/bin/bash
# aligner search kernel: threads per utterance (KHG_ALIGN_NT) x frames per likelihood tile (KHG_ALIGN_FORCE_FC) at C5
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4f_*
for cfg in "128 32" "64 32" "32 32" "32 16" "32 8" "64 16" "64 8"; do
  set -- $cfg
  echo "## NT=$1 FC=$2" >> $O/r4f_align_nt.txt
  KHG_ALIGN_NT=$1 KHG_ALIGN_FORCE_FC=$2 KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 4 --check 4 2>&1 | grep -v "^khg_align_batch host" | cut -c1-420 | tail -4 >> $O/r4f_align_nt.txt
done
grep -o "## NT.*\|search [0-9.]* ms\|\"value_device_feats\": [0-9.]*" $O/r4f_align_nt.txt | paste -s -d' ' | sed 's/## /\n/g'

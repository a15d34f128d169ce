#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4l_*
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_em_recipe.py -m gpu -q -x 2>&1 | tail -3 > $O/r4l_tests.txt
cat $O/r4l_tests.txt
for cfg in "0 0" "128 32"; do
  set -- $cfg
  echo "## NT=$1 FC=$2 (0 = chosen)" >> $O/r4l_align.txt
  if [ "$1" = "0" ]; then KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 4 --check 4 2>&1 | grep -v "^khg_align_batch host" | cut -c1-420 | tail -3 >> $O/r4l_align.txt
  else KHG_ALIGN_NT=$1 KHG_ALIGN_FORCE_FC=$2 KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 4 --check 4 2>&1 | grep -v "^khg_align_batch host" | cut -c1-420 | tail -3 >> $O/r4l_align.txt; fi
done
grep -o "## NT.*\|FC [0-9]* NT [0-9]*\|search [0-9.]* ms\|\"value_device_feats\": [0-9.]*" $O/r4l_align.txt | paste -s -d' ' | sed 's/## /\n/g'

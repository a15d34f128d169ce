#!/bin/bash
# aligner dense block: tile lists per frame tile (plain launch) against per pair of tiles (CTA pairs) at C5
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4n_*
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_em_recipe.py -m gpu -q -x 2>&1 | tail -3 > $O/r4n_tests.txt
cat $O/r4n_tests.txt
for r in 1 2; do for sh in 0; do
  echo "## SUBSET_SHIFT=$sh" >> $O/r4n_align.txt
  KHG_ALIGN_SUBSET_SHIFT=$sh KHG_ALIGN_TIMING=1 timeout 300 python tools/bench_align.py --reps 4 --check 4 2>&1 | grep -v "^khg_align_batch host" | cut -c1-420 | tail -3 >> $O/r4n_align.txt
done; done
grep -o "## SUBSET.*\|dense [0-9.]* ms ([0-9.]* %\|search [0-9.]* ms\|\"value_device_feats\": [0-9.]*" $O/r4n_align.txt | paste -s -d' ' | sed 's/## /\n/g'

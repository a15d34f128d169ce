#!/bin/bash
# compute-sanitizer over what changed since r2_sanitizer: K3t's 256-bit / 16-byte-chunk row loads and MN-major posterior tile,
# khg_estep's grouped statistics (host halves), the aligner's launch shapes (64 threads x 8-frame tile)
cd "$(dirname "$0")/.."
O=gpurun_out
F=$O/r4k_sanitizer.txt
rm -f $F
echo "## memcheck: tests/test_gpu_stats_tc.py (dims 1..40: aligned 256-bit, 128-bit, 16-byte-chunk and scalar row loads)" >> $F
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_stats_tc.py -q -x 2>&1 | tail -5 >> $F
echo "## memcheck: khg_estep statistics groups + device / host paths" >> $F
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "estep" 2>&1 | tail -5 >> $F
echo "## memcheck: aligner (all launch shapes the tests reach; KHG_ALIGN_NT=64 KHG_ALIGN_FORCE_FC=8 forced on the realistic cases)" >> $F
KHG_ALIGN_NT=64 KHG_ALIGN_FORCE_FC=8 timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_align.py -q -x -k "matches_oracle or reference_rule or tiles_the_graphs_need or min_active" 2>&1 | tail -5 >> $F
echo "## racecheck: K3t (dim 39: chunk loads) and the aligner at 64 threads" >> $F
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_stats_tc.py -q -x -k "vs_oracle and 39" 2>&1 | tail -5 >> $F
KHG_ALIGN_NT=64 KHG_ALIGN_FORCE_FC=8 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_align.py -q -x -k "tiles_the_graphs_need" 2>&1 | tail -5 >> $F
cat $F

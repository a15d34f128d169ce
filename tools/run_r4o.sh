#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_em_recipe.py -m gpu -q -x 2>&1 | tail -3 > $O/r4o_tests.txt
cat $O/r4o_tests.txt

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4x_*
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_em_recipe.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 > $O/r4x_tests.txt
cat $O/r4x_tests.txt
for r in 1 2; do timeout 300 python tools/bench_align.py --reps 4 --check 4 2>&1 | tail -1 | grep -o '"value_device_feats": [0-9.]*\|"e2e_host_feats": [0-9.]*' | paste -s -d' ' | tee -a $O/r4x_align.txt; done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > $O/r4x_bench.json 2> $O/r4x_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4x_bench.json').read().strip().splitlines()[-1])
a=d['workloads']['align_c5']; print(a['value'], a['value_first_alignment'], a['e2e_host_feats']['value'])
PY

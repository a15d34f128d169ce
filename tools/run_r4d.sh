#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r4d_*
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_em_recipe.py tests/test_reference_api.py -m gpu -q -x 2>&1 | tail -5 > $O/r4d_tests.txt
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r4d_bench_n1.json 2> $O/r4d_bench_n1.err
cat $O/r4d_tests.txt; tail -c 300 $O/r4d_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4d_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['parity_check']['ok'], d['roofline']['frac'], d['roofline']['dense_frames_per_s'], d['roofline']['share_of_step'])
print({k:(round(v.get('value',0)), v.get('roofline',{}).get('frac')) for k,v in d['workloads'].items() if isinstance(v,dict)})
print(d['workloads']['w_aligned_c4'])
PY

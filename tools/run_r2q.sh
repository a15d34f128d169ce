#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2q_gputests.txt
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r2q_bench_n1.json 2> $O/r2q_bench_n1.err
tail -8 $O/r2q_gputests.txt; tail -c 300 $O/r2q_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['parity_check'])
print({k:(v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d['workloads'].items() if isinstance(v,dict)})
print({k:v for k,v in d.items() if 'stats' in k})
PY

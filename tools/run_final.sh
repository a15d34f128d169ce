#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/r5l_gputests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r5l_smoke.txt 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r5l_bench_n1.json 2> $O/r5l_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/r5l_bench_ref.json 2> $O/r5l_bench_ref.err
tail -4 $O/r5l_gputests.txt; tail -2 $O/r5l_smoke.txt; tail -c 300 $O/r5l_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5l_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['parity_check']['ok'], d['roofline']['frac'], d['roofline']['frac_of_tf32_peak_in_run'])
print({k:(round(v.get('value',0)), v.get('roofline',{}).get('frac')) for k,v in d['workloads'].items() if isinstance(v,dict)})
print(d['workloads']['align_c5'].get('dense_tile_units_computed_rank0'), d['workloads']['align_c5']['e2e_host_feats']['value'])
r=json.loads(open('gpurun_out/r5l_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['cores'])
PY

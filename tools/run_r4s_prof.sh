#!/bin/bash
# round-2 final profiles: ncu --set full of the dense kernel (K1, fp16 split), of the statistics kernel (K3t) and of the search kernel;
# launch list of a short bench run (gpu__time_duration per launch)
cd "$(dirname "$0")/.."
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:loglikes_tc -s 2 -c 1 -f -o $O/r4s_k1 python tools/prof_dense.py > $O/r4s_ncu_k1.log 2>&1
python tools/ncu_summary.py $O/r4s_k1.ncu-rep > $O/r4s_ncu_loglikes_tc_f16.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:stats_tc_kernel -s 1 -c 1 -f -o $O/r4s_k3t python tools/prof_dense.py 2000000 > $O/r4s_ncu_k3t.log 2>&1
python tools/ncu_summary.py $O/r4s_k3t.ncu-rep > $O/r4s_ncu_stats_tc.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:viterbi_kernel -s 1 -c 1 -f -o $O/r4s_vit python tools/bench_align.py --reps 1 --check 0 > $O/r4s_ncu_vit.log 2>&1
python tools/ncu_summary.py $O/r4s_vit.ncu-rep > $O/r4s_ncu_viterbi.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r4s_launches.csv python bench.py --frames 3000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-workloads > $O/r4s_ncu_launches.log 2>&1
tail -3 $O/r4s_ncu_loglikes_tc_f16.txt; grep -c . $O/r4s_launches.csv

# compute-sanitizer over the aligner / mix-up / M-step / Gaussian-selection parity tests of the final build
out=gpurun_out/r1y_sanitizer.txt
echo "# compute-sanitizer on the final build (aligner with redux.sync reductions, mix-up / mix-down, M-step, Gaussian selection)" > $out
echo "## memcheck" >> $out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_align.py tests/test_mixup.py tests/test_gpu_mstep.py tests/test_gselect.py -m gpu -q -x 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -20 >> $out
echo "## racecheck (aligner)" >> $out
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_align.py -m gpu -q -x 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard" | head -20 >> $out
echo "## synccheck (aligner)" >> $out
timeout 300 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_align.py -m gpu -q -x 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Barrier" | head -20 >> $out
cat $out

#!/bin/bash
# compute-sanitizer over a small slice of the GPU tests (hand-rolled mbarrier / TMEM / cluster pipelines).
# usage: tools/gpu_r02_sanitize.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
SEL='small_batch or variable_length or golden_and_properties or paired_distance or (gemm and 515) or (gemm and 999) or (attention and lengths0) or (loss_value and 0.1) or score_entry'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "$SEL" > $out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/${tag}_sanitizer_${tool}.log | tail -3
done

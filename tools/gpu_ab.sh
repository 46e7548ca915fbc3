#!/bin/bash
# A/B of an environment switch on the bench workload in ONE box (per-process placement differs by up to +-4 %):
# usage: tools/gpu_ab.sh <tag> <ENVVAR> [valueA valueB]; alternates step_trace runs A B A B
tag=$1; var=$2; va=${3:-0}; vb=${4:-1}
out=gpurun_out; mkdir -p $out
for rep in 1 2; do
  for v in $va $vb; do
    env $var=$v timeout 300 python tools/step_trace.py > $out/${tag}_trace_${var}_${v}_$rep.log 2>&1
    echo "== $var=$v rep $rep"; grep -E "span|gemm " $out/${tag}_trace_${var}_${v}_$rep.log
  done
done

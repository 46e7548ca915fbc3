#!/bin/bash
# A/B of one environment switch: parity file (achieved errors) per value + alternating step traces.  usage: tools/gpu_env_ab.sh <tag> <VAR> <a> <b>
tag=$1; var=$2; va=$3; vb=$4
out=gpurun_out; mkdir -p $out
for v in $va $vb; do
  rm -f $out/parity_achieved.jsonl
  env $var=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_parity_${var}_$v.log 2>&1; echo "parity $var=$v: $(tail -1 $out/${tag}_parity_${var}_$v.log)"
  cp $out/parity_achieved.jsonl $out/${tag}_achieved_${var}_$v.jsonl 2>/dev/null
done
for rep in 1 2; do
  for v in $va $vb; do
    env $var=$v timeout 300 python tools/step_trace.py > $out/${tag}_trace_${var}_${v}_$rep.log 2>&1
    echo "== $var=$v rep $rep: $(grep span $out/${tag}_trace_${var}_${v}_$rep.log) | $(grep -E 'gemm (conv1|fc1|fc2)' $out/${tag}_trace_${var}_${v}_$rep.log | tr '\n' ' ')"
  done
done

#!/bin/bash
# A/B of two builds of the library on the bench step in ONE box: usage tools/gpu_ab_lib.sh <tag> <libA.so> <libB.so>
tag=$1; la=$2; lb=$3
out=gpurun_out; mkdir -p $out
for rep in 1 2; do
  for l in $la $lb; do
    NOMAD_B200_LIB=$PWD/$l timeout 300 python tools/step_trace.py > $out/${tag}_trace_$(basename $l .so)_$rep.log 2>&1
    echo "== $l rep $rep"; grep -E "span|conv0_mma|attention_fa|gemm conv1|gemm fc1|gemm qkv|gemm out|gemm fc2" $out/${tag}_trace_$(basename $l .so)_$rep.log
  done
done

#!/bin/bash
# conv0 variants: 1 = row-major tiles + quad shuffles, 2 / 3 = channel-permuted groups of 2 / 4 tiles, 4 / 5 = 2 + shared A fragments (4 / 3 blocks per SM)
out=gpurun_out; mkdir -p $out; tag=${1:-r02r}; shift
vars=${@:-"2 4 5"}
for v in $vars; do
  [ $v = 2 ] && continue
  NOMAD_B200_CONV0_MMA=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_parity_conv0_$v.log 2>&1; echo "parity v=$v: $(tail -1 $out/${tag}_parity_conv0_$v.log)"
done
for rep in 1 2; do
  for v in $vars; do
    NOMAD_B200_CONV0_MMA=$v timeout 300 python tools/step_trace.py > $out/${tag}_trace_conv0_${v}_$rep.log 2>&1
    echo "== conv0 variant $v rep $rep: $(grep span $out/${tag}_trace_conv0_${v}_$rep.log) | $(grep conv0 $out/${tag}_trace_conv0_${v}_$rep.log | head -1)"
  done
done

#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r02v}
for v in 0 1; do for c in out_rf out_rfh big768 resid; do NOMAD_B200_RESID_TMA=$v python tools/probe_gemm.py $c 2>&1 | sed "s/^/RTMA=$v /"; done; done | tee $out/${tag}_rtma_probe.log
for rep in 1 2; do
  for v in 0 1; do
    NOMAD_B200_CONV0_MMA=2 NOMAD_B200_RESID_TMA=$v timeout 300 python tools/step_trace.py > $out/${tag}_trace_rtma_${v}_$rep.log 2>&1
    echo "== RTMA=$v rep $rep: $(grep span $out/${tag}_trace_rtma_${v}_$rep.log) | $(grep -E 'gemm (out|fc2)' $out/${tag}_trace_rtma_${v}_$rep.log | tr '\n' ' ')"
  done
done
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 67 -c 1 -o $out/${tag}_ncu_outproj1 python tools/ncu_target.py 2 > /dev/null 2>&1
ls $out | grep ${tag}

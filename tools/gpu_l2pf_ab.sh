#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r02t}
for v in 0 1 2; do for c in out_rf out_rfh big768; do NOMAD_B200_RESID_L2PF=$v python tools/probe_gemm.py $c 2>&1 | sed "s/^/L2PF=$v /"; done; done | tee $out/${tag}_l2pf_probe.log
for rep in 1 2; do
  for v in 0 1 2; do
    NOMAD_B200_CONV0_MMA=2 NOMAD_B200_RESID_L2PF=$v timeout 300 python tools/step_trace.py > $out/${tag}_trace_l2pf_${v}_$rep.log 2>&1
    echo "== L2PF=$v rep $rep: $(grep span $out/${tag}_trace_l2pf_${v}_$rep.log) | $(grep -E 'gemm (out|fc2)' $out/${tag}_trace_l2pf_${v}_$rep.log | tr '\n' ' ')"
  done
done

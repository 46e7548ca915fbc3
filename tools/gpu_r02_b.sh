#!/bin/bash
# round-2 check after the loss graph / ingest / fp32-mode work.  usage: tools/gpu_r02_b.sh <tag>
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
export NOMAD_B200_PARITY_LOG=$out/${tag}_parity_achieved.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -8 $out/${tag}_pytest_gpu.log
timeout 300 python tools/loss_trace.py > $out/${tag}_loss_trace.log 2>&1; head -12 $out/${tag}_loss_trace.log | grep -v Warn
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err
TAG=$tag python - <<'P'
import json,sys
d=json.load(open("gpurun_out/" + __import__("os").environ.get("TAG", "r02i") + "_bench.json"))
for k in ("value","ms_per_step"): print(k, d[k])
for k in ("e2e","roofline","pairwise","loss","fp32_mode","parity","sharded_c3"):
    print(k, json.dumps({x:y for x,y in (d[k] or {}).items() if x not in ("workload","timed","roof","api","kernel","peak_source","traffic_source","note","flops_definition")}))
P
timeout 600 python tools/bench_files.py 3000 300 > $out/${tag}_bench_files.json 2>> $out/${tag}_bench.err; cat $out/${tag}_bench_files.json

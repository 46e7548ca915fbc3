#!/bin/bash
# round-2 first check: tests, smoke, bench (ours / reference / library).  usage: tools/gpu_r02_a.sh <tag>
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
export NOMAD_B200_PARITY_LOG=$out/${tag}_parity_achieved.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -15 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err; cut -c1-3000 $out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err; cut -c1-200 $out/${tag}_bench_reference.json
timeout 300 python bench.py --impl library --steps 5 --warmup 3 > $out/${tag}_bench_library.json 2>> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench_library.json

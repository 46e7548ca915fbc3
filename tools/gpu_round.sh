#!/bin/bash
# One GPU round of evidence for profiles/: tests, smoke, bench, launch list, ncu captures.  usage: tools/gpu_round.sh <tag>
# (run under gpurun from the repo root; writes gpurun_out/<tag>_*)
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cut -c1-400 $out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench_reference.json
timeout 300 python tools/step_trace.py > $out/${tag}_step_trace.log 2>&1; grep "span" $out/${tag}_step_trace.log
# every launch of one step (76 kernels per step; 2 warm-up steps skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -s 152 -c 76 --csv --log-file $out/${tag}_launches.csv python tools/ncu_target.py 3 > /dev/null 2>&1
# DRAM bytes of every GEMM launch of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc -s 55 -c 55 --csv --log-file $out/${tag}_gemm_dram.csv python tools/ncu_target.py 2 > /dev/null 2>&1
# full captures: second step's conv1, qkv0, out0, fc1_0, fc2_0 (GEMM launch indices 55+0, +7, +8, +9, +10)
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 55 -c 1 -o $out/${tag}_ncu_gemm_conv1 python tools/ncu_target.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 62 -c 4 -o $out/${tag}_ncu_gemm_layer0 python tools/ncu_target.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attention_fa|conv0_mma|posconv_kernel" -s 3 -c 3 -o $out/${tag}_ncu_misc python tools/ncu_target.py 2 > /dev/null 2>&1
timeout 900 python tools/bench_configs.py > $out/${tag}_bench_configs.jsonl 2> $out/${tag}_bench_configs.err; cut -c1-260 $out/${tag}_bench_configs.jsonl
timeout 300 python tools/probe_cdist.py 2>&1 | grep tcgen05 > $out/${tag}_cdist_probe.log; tail -4 $out/${tag}_cdist_probe.log
ls -la $out | grep ${tag}_ | awk '{print $5, $9}'

#!/bin/bash
# multi-GPU evidence: the N-rank NCCL tests + bench.py under torchrun.  usage: tools/gpu_r02_dist.sh <tag> <ngpus>
tag=${1:-r02}
n=${2:-2}
out=gpurun_out
mkdir -p $out
export NOMAD_B200_PARITY_LOG=$out/${tag}_parity_achieved_n${n}.jsonl
nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_round2.py -m gpu -q -k "dist or sharded or bit_identical" > $out/${tag}_pytest_gpu_dist_n${n}.log 2>&1; tail -5 $out/${tag}_pytest_gpu_dist_n${n}.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
tail -3 $out/${tag}_bench_n${n}.err | cut -c1-300; cat $out/${tag}_bench_n${n}.json | cut -c1-200
grep -m3 -i "nvls\|P2P/CUMEM\|via P2P" $out/${tag}_bench_n${n}.err | cut -c1-200

#!/bin/bash
# 8-GPU box visit: bench at 8 GPUs (exchange and recompute mode) and at 4 GPUs
tag=${1:-run8}
for n in 8 4; do
  NZ_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err; echo "bench$n rc=$?"
  grep "ms_step" gpurun_out/${tag}_bench${n}.err | head -3
done
NZ_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 --mode recompute --no-e2e --no-configs > gpurun_out/${tag}_bench8_recompute.json 2> gpurun_out/${tag}_bench8_recompute.err; echo "recompute rc=$?"
grep "ms_step" gpurun_out/${tag}_bench8_recompute.err | head -2

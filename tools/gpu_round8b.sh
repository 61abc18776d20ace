#!/bin/bash
tag=${1:-run8}
n=$(nvidia-smi -L | wc -l)
NZ_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err; echo "bench$n rc=$?"
grep "ms_step" gpurun_out/${tag}_bench${n}.err | head -8
NZ_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 --mode recompute --no-e2e --no-configs > gpurun_out/${tag}_bench${n}_recompute.json 2> gpurun_out/${tag}_bench${n}_recompute.err; echo "recompute rc=$?"
grep "ms_step" gpurun_out/${tag}_bench${n}_recompute.err | head -2

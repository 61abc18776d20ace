#!/bin/bash
# One GPU-box visit: tests, 1-GPU bench, N-GPU bench (N = visible GPUs).  Everything lands in gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench1.json 2> gpurun_out/${tag}_bench1.err; echo "bench1 rc=$?"
if [ "$n" -gt 1 ]; then
  NZ_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err; echo "bench$n rc=$?"
  tail -3 gpurun_out/${tag}_bench${n}.err
fi
tail -3 gpurun_out/${tag}_bench1.err
if [ -n "$NZ_ROUND_REFERENCE" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err; echo "reference rc=$?"
fi

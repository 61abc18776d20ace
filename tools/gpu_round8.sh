#!/bin/bash
# 8-GPU visit: the N-GPU bench (both engines) and the D2H probe.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-run8}
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
NZ_BENCH_VERBOSE=1 run 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err; echo "bench$n lib rc=$?"
grep "ms_step" gpurun_out/${tag}_bench${n}.err | head -8
NZ_BENCH_VERBOSE=1 run 29512 bench.py --gpus $n --steps 20 --warmup 5 --engine python --no-e2e --no-configs > gpurun_out/${tag}_bench${n}_py.json 2> gpurun_out/${tag}_bench${n}_py.err; echo "bench$n python rc=$?"
grep "ms_step" gpurun_out/${tag}_bench${n}_py.err | head -3
run 29513 tools/d2h_probe_multi.py 2 > gpurun_out/${tag}_d2h_bound.txt 2> gpurun_out/${tag}_d2h_bound.err; echo "probe bound rc=$?"
NZ_PROBE_NO_BIND=1 run 29514 tools/d2h_probe_multi.py 2 > gpurun_out/${tag}_d2h_unbound.txt 2> gpurun_out/${tag}_d2h_unbound.err; echo "probe unbound rc=$?"
cat gpurun_out/${tag}_d2h_bound.txt gpurun_out/${tag}_d2h_unbound.txt
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/${tag}_topo.txt; free -g >> gpurun_out/${tag}_topo.txt

#!/bin/bash
# Final ncu evidence of a round: launch list of the bench chain, then ONE --set full capture of every kernel of a chain step
# (raw page exported on the box; the .ncu-rep stays there).  Numbers printed by a run under ncu are never bench values.
tag=${1:-r2final}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs"
$B > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
B1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs"
ncu --set full --clock-control none -c 40 -o /tmp/${tag}_chain $B1 > gpurun_out/${tag}_ncu_chain.log 2>&1
echo "full capture rc=$?"
ncu -i /tmp/${tag}_chain.ncu-rep --page raw --csv > gpurun_out/${tag}_chain.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}_chain.raw.csv > gpurun_out/${tag}_ncu_chain_summary.txt 2>&1
ls -la gpurun_out/${tag}_*

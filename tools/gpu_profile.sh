#!/bin/bash
# ncu evidence of the round: launch list of the bench chain + --set full captures of the kernels that changed.
tag=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs"
$B > gpurun_out/${tag}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/flow_only.py 16384 psr > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fbm_psr_pair -c 1 -o gpurun_out/${tag}_psr python tools/flow_only.py 16384 psr > gpurun_out/${tag}_ncu_psr.log 2>&1
echo "psr rc=$?"
NZ_FLOW_GROUP=4 python tools/flow_reg_only.py 16384 > /dev/null 2>&1 && NZ_FLOW_GROUP=4 ncu --set full --clock-control none --import-source on -k regex:flow_group_kernel -c 1 -o gpurun_out/${tag}_flow_group python tools/flow_reg_only.py 16384 > gpurun_out/${tag}_ncu_flow_group.log 2>&1
echo "flow group rc=$?"
python tools/flow_only.py 16384 noise > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fbm_simplex_pair -c 1 -o gpurun_out/${tag}_fbm python tools/flow_only.py 16384 noise > gpurun_out/${tag}_ncu_fbm.log 2>&1
echo "fbm rc=$?"
for f in psr flow_group fbm; do ncu -i gpurun_out/${tag}_$f.ncu-rep --page raw --csv > gpurun_out/${tag}_$f.raw.csv 2>/dev/null; done
ls -la gpurun_out/${tag}_* | head -20

#!/bin/bash
# Round-2 GPU run R (1 GPU): ncu --set full of K2 (symmetric forward) and K3 at cfg4 and cfg2, final code.
mkdir -p gpurun_out
T=${1:-R}
CMD4="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-extras"
$CMD4 > gpurun_out/${T}_plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:infonce_tc -s 6 -c 2 -o gpurun_out/${T}_ncu_cfg4 -f $CMD4 > gpurun_out/${T}_ncu_full_cfg4.log 2>&1
echo "ncu cfg4 rc=$?"
CMD="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/${T}_plain_cfg2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:infonce_tc -s 6 -c 2 -o gpurun_out/${T}_ncu_cfg2 -f $CMD > gpurun_out/${T}_ncu_full_cfg2.log 2>&1
echo "ncu cfg2 rc=$?"
ls -la gpurun_out/${T}_*

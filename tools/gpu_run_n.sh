#!/bin/bash
# Round-2 GPU run N (1 GPU): ncu --set full of the symmetric forward (n=8192 x 256 and cfg2) with source-level sampling.
mkdir -p gpurun_out
T=${1:-N}
CMD="python tools/sym_debug.py 8192 256"
$CMD > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwdsym -c 1 -o gpurun_out/${T}_sym_d256 -f $CMD > gpurun_out/${T}_ncu1.log 2>&1
echo "ncu d256 rc=$?"
CMD="python tools/sym_debug.py 4096 128"
$CMD >> gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwdsym -c 1 -o gpurun_out/${T}_sym_d128 -f $CMD > gpurun_out/${T}_ncu2.log 2>&1
echo "ncu d128 rc=$?"
ls -la gpurun_out/${T}_*

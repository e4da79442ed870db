#!/bin/bash
# Round-2 GPU run H (1 GPU): default bench + ncu --set full of the tensor-core kernels at cfg2 and cfg4.
mkdir -p gpurun_out
T=${1:-H}
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
CMD="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/${T}_plain_cfg2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:infonce_tc -s 6 -c 4 -o gpurun_out/${T}_ncu_cfg2 $CMD > gpurun_out/${T}_ncu_full_cfg2.log 2>&1
echo "ncu cfg2 rc=$?"
CMD4="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-extras"
$CMD4 > gpurun_out/${T}_plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none -k regex:infonce_tc -s 6 -c 2 -o gpurun_out/${T}_ncu_cfg4 $CMD4 > gpurun_out/${T}_ncu_full_cfg4.log 2>&1
echo "ncu cfg4 rc=$?"
$CMD4 > gpurun_out/${T}_plain_cfg4b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_cfg4.csv $CMD4 > gpurun_out/${T}_ncu_launch4.log 2>&1
echo "ncu launches cfg4 rc=$?"
ls -la gpurun_out | tail -8

#!/bin/bash
# Round-2 GPU run P (1 GPU): full GPU test suite + default bench (with extras) + cfg2 bench + ncu launch lists.
mkdir -p gpurun_out
T=${1:-P}
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -25 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench.err
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_cfg2.json 2>/dev/null
CMD="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/${T}_plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/${T}_launches_cfg2.csv $CMD > gpurun_out/${T}_ncu_launch2.log 2>&1
echo "ncu launches cfg2 rc=$?"
CMD4="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-extras"
$CMD4 > gpurun_out/${T}_plain_cfg4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_cfg4.csv $CMD4 > gpurun_out/${T}_ncu_launch4.log 2>&1
echo "ncu launches cfg4 rc=$?"
ls -la gpurun_out | tail -8

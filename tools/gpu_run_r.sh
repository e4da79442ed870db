#!/bin/bash
# Round-2 GPU run R (1 GPU): ncu --set full of the symmetric forward (K2) at cfg4 and cfg2, final code; the summaries are
# made on the box (the reports together exceed what gpurun copies back), plus an A/B of the shared-memory carve-out.
mkdir -p gpurun_out
T=${1:-R}
CMD4="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-extras"
$CMD4 > gpurun_out/${T}_plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwdsym -s 3 -c 1 -o gpurun_out/${T}_ncu_cfg4 -f $CMD4 > gpurun_out/${T}_ncu_full_cfg4.log 2>&1
echo "ncu cfg4 rc=$?"
python tools/ncu_summary.py full gpurun_out/${T}_ncu_cfg4.ncu-rep > gpurun_out/${T}_ncu_full_sym_cfg4.txt 2>&1
CMD="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/${T}_plain_cfg2.log 2>&1 && \
ncu --set full --clock-control none -k regex:fwdsym -s 3 -c 1 -o gpurun_out/${T}_ncu_cfg2 -f $CMD > gpurun_out/${T}_ncu_full_cfg2.log 2>&1
echo "ncu cfg2 rc=$?"
python tools/ncu_summary.py full gpurun_out/${T}_ncu_cfg2.ncu-rep > gpurun_out/${T}_ncu_full_sym_cfg2.txt 2>&1
rm -f gpurun_out/${T}_ncu_cfg2.ncu-rep
for C in 1 0 1 0; do
SM3_CARVEOUT=$C timeout 300 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-extras > gpurun_out/${T}_cfg2_carve$C.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_cfg2_carve$C.json').read().strip().splitlines()[-1])
print('carveout=$C', 'eager ms', round(d['ms_per_step'],4), 'graph ms', round(d['cuda_graph']['ms_per_step'],4), d['stages_ms'])
PY
done
ls -la gpurun_out/${T}_*

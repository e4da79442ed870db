#!/bin/bash
# Round-2 GPU run T (1 GPU): packed-math backward kernels -- parity tests, then cfg4 / cfg2 benches.
mkdir -p gpurun_out
T=${1:-T}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 -k "tensor_core or backward_forms or cfg2_full or known_answers or fp16_inputs or golden or symmetric" -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/${T}_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_$i.json 2>/dev/null
timeout 300 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-extras > gpurun_out/${T}_bench_cfg2_$i.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/T_bench*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['ms_per_step'],4), d['stages_ms'], r['parity']['ok'], r['parity']['loss_relerr'], r['parity']['grad_relerr_rowblock'], (d.get('cuda_graph') or {}).get('ms_per_step'))
PY
CMD="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/${T}_plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/${T}_launches_cfg2.csv $CMD > gpurun_out/${T}_ncu_launch2.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${T}_launches_cfg2.csv | head -9 | cut -c1-120

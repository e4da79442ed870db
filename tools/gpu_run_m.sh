#!/bin/bash
# Round-2 GPU run M (1 GPU): symmetric forward variants -- parity tests, trace, POLY x NQ sweep at cfg4 / cfg2.
mkdir -p gpurun_out
T=${1:-M}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "symmetric_forward or cfg2_full or known_answers" -p no:cacheprovider > gpurun_out/${T}_pytest_sym.log 2>&1
echo "pytest sym rc=$?"; tail -8 gpurun_out/${T}_pytest_sym.log
SM3_LIB_PATH=skin_sm3_b200/lib/libsm3_b200_trace.so TRACE_BWD_VERSIONS=4 timeout 300 python tools/trace_cfg2.py 4096 128 2>&1 | tail -14 > gpurun_out/${T}_trace_cfg2.txt
head -12 gpurun_out/${T}_trace_cfg2.txt
SM3_LIB_PATH=skin_sm3_b200/lib/libsm3_b200_trace.so TRACE_BWD_VERSIONS=1 timeout 300 python tools/trace_cfg2.py 8192 256 2>&1 | tail -20 > gpurun_out/${T}_trace_n8192_d256.txt
cat gpurun_out/${T}_trace_n8192_d256.txt
for Q in 4 2; do for P in 0 2 4; do
SM3_TC_SYM_NQ=$Q SM3_TC_POLY=$P timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_q${Q}_poly$P.json 2> gpurun_out/${T}_bench.err
echo "bench q$Q poly$P rc=$?"
SM3_TC_SYM_NQ=$Q SM3_TC_POLY=$P timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_cfg2_q${Q}_poly$P.json 2>/dev/null
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench*_q*poly*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],4), r['stages_ms'], r['parity']['ok'], r['parity']['loss_relerr'], (d.get('cuda_graph') or {}).get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY

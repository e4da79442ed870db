#!/bin/bash
# Round-2 GPU run L (1 GPU): trace of the symmetric forward, tiled top-k tests + timing, sym tests.
mkdir -p gpurun_out
T=${1:-L}
SM3_LIB_PATH=skin_sm3_b200/lib/libsm3_b200_trace.so TRACE_BWD_VERSIONS=4 timeout 300 python tools/trace_cfg2.py 4096 128 > gpurun_out/${T}_trace_cfg2.txt 2>&1
echo "trace rc=$?"; tail -32 gpurun_out/${T}_trace_cfg2.txt
SM3_LIB_PATH=skin_sm3_b200/lib/libsm3_b200_trace.so TRACE_BWD_VERSIONS=1 timeout 300 python tools/trace_cfg2.py 8192 256 > gpurun_out/${T}_trace_n8192_d256.txt 2>&1
tail -24 gpurun_out/${T}_trace_n8192_d256.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 -k "topk or knn or retrieval or symmetric_forward or cluster_memory" -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --extras-child > gpurun_out/${T}_extras.json 2> gpurun_out/${T}_extras.err
echo "extras rc=$?"; tail -c 600 gpurun_out/${T}_extras.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/'+__import__('sys').argv[1] if False else 'gpurun_out/L_extras.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('retrieval')), json.dumps(d.get('small_shapes'))[:600])
PY

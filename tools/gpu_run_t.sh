#!/bin/bash
# Round-2 GPU run T (1 GPU): quick check of a kernel change -- the parity tests that touch it, then the retrieval probe.
mkdir -p gpurun_out
T=${1:-T}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 -k "topk or knn or retrieval or cluster_memory" -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/${T}_pytest.log
timeout 300 python - <<'PY' > gpurun_out/${T}_retrieval.json 2> gpurun_out/${T}_retrieval.err
import json, bench, skin_sm3_b200 as sm3, torch
torch.cuda.set_device(0)
print(json.dumps(bench.retrieval_probe(sm3)))
PY
cat gpurun_out/${T}_retrieval.json; tail -3 gpurun_out/${T}_retrieval.err

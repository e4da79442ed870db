#!/bin/bash
# Round-2 GPU run Q (1 GPU): cross-rank symmetric forward played on one GPU, prototype heads, head-block probe.
mkdir -p gpurun_out
T=${1:-Q}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 -k "cross_rank_symmetric or proto_heads or symmetric_forward" -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/${T}_pytest.log
timeout 300 python - <<'PY' > gpurun_out/${T}_headblock.json 2> gpurun_out/${T}_headblock.err
import json, bench, skin_sm3_b200 as sm3, torch
torch.cuda.set_device(0)
print(json.dumps(bench.head_block_probe(sm3)))
PY
cat gpurun_out/${T}_headblock.json; tail -3 gpurun_out/${T}_headblock.err

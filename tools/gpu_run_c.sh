#!/bin/bash
# Round-2 GPU run C (1 GPU, short): parity sanity of the 128-column backward, per-variant kernel times, pipeline traces.
mkdir -p gpurun_out
T=D
python - > gpurun_out/${T}_sanity.log 2>&1 <<'PY'
import os, torch, numpy as np, sys
sys.path.insert(0, ".")
import skin_sm3_b200 as sm3
from oracle import sm3_oracle as O
ok = True
for v, poly in (("4", "0"), ("4", "2"), ("3", "2")):
    os.environ["SM3_TC_BWD_V"] = v; os.environ["SM3_TC_BWD_POLY"] = poly
    sm3.reload_env()
    for n, d in ((700, 128), (320, 64), (2100, 128), (64, 128), (4096, 128)):
        g = torch.Generator().manual_seed(n)
        p1 = torch.randn(n, d, generator=g).bfloat16(); p2 = (p1.float() + 0.5 * torch.randn(n, d, generator=g)).bfloat16()
        a, b = p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(True)
        loss = sm3.fused_infonce(a, b, 0.1, precision="bf16"); loss.backward(); torch.cuda.synchronize()
        ref, r1, r2 = O.infonce_closed_form(p1.float().numpy(), p2.float().numpy(), 0.1, chunk=2048)
        e = max(np.abs(a.grad.float().cpu().numpy() - r1).max() / np.abs(r1).max(), np.abs(b.grad.float().cpu().numpy() - r2).max() / np.abs(r2).max())
        print(f"bwd_v={v} poly={poly} n={n} d={d} loss={loss.item():.5f} ref={ref:.5f} grad_relerr={e:.3e}", flush=True)
        ok &= (abs(loss.item() - ref) < 2e-2 * abs(ref)) and e < 2e-2
print("SANITY_OK" if ok else "SANITY_FAIL")
PY
tail -4 gpurun_out/${T}_sanity.log
python - > gpurun_out/${T}_variants.json 2> gpurun_out/${T}_variants.err <<'PY'
import sys, json
sys.path.insert(0, ".")
import importlib.util, torch
spec = importlib.util.spec_from_file_location("bench", "bench.py"); b = importlib.util.module_from_spec(spec); sys.argv = ["x"]; spec.loader.exec_module(b)
import skin_sm3_b200 as sm3
torch.cuda.set_device(0)
print(json.dumps(b.tc_kernel_probe(sm3), indent=1))
PY
cat gpurun_out/${T}_variants.json | head -60
python tools/umma_rate.py > gpurun_out/${T}_umma_rate.txt 2>&1
export SM3_LIB_PATH=skin_sm3_b200/lib/libsm3_b200_trace.so
TRACE_BWD_VERSIONS=4 python tools/trace_cfg2.py 4096 128 > gpurun_out/${T}_trace_cfg2.txt 2>&1
tail -3 gpurun_out/${T}_trace_cfg2.txt

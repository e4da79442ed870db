#!/bin/bash
# Round-2 GPU run A (1 GPU): sanity of the new backward kernel, full GPU test suite, default bench, cfg2 ncu captures.
# Everything lands in gpurun_out/.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_run_a.sh'
mkdir -p gpurun_out
T=A
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/${T}_clocks.csv 2>/dev/null &
SMI=$!
python - > gpurun_out/${T}_sanity.log 2>&1 <<'PY'
import os, torch, numpy as np, sys
sys.path.insert(0, ".")
import skin_sm3_b200 as sm3
from oracle import sm3_oracle as O
ok = True
for v in ("1", "2"):
    os.environ["SM3_TC_BWD_V"] = v
    sm3.lib().sm3_debug_reload_env()
    for n, d in ((700, 128), (320, 64), (1536, 256), (200, 192)):
        g = torch.Generator().manual_seed(n)
        p1 = torch.randn(n, d, generator=g).bfloat16(); p2 = (p1.float() + 0.5 * torch.randn(n, d, generator=g)).bfloat16()
        a, b = p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(True)
        loss = sm3.fused_infonce(a, b, 0.1, precision="bf16"); loss.backward(); torch.cuda.synchronize()
        ref, r1, r2 = O.infonce_closed_form(p1.float().numpy(), p2.float().numpy(), 0.1)
        e = np.abs(a.grad.float().cpu().numpy() - r1).max() / np.abs(r1).max()
        print(f"bwd_v={v} n={n} d={d} loss={loss.item():.5f} ref={ref:.5f} grad_relerr={e:.3e}", flush=True)
        ok &= (abs(loss.item() - ref) < 2e-2 * abs(ref)) and e < 2e-2
print("SANITY_OK" if ok else "SANITY_FAIL")
PY
tail -3 gpurun_out/${T}_sanity.log
if ! grep -q SANITY_OK gpurun_out/${T}_sanity.log; then
  echo "new backward kernel failed sanity: falling back to SM3_TC_BWD_V=1 for the rest of this run" | tee -a gpurun_out/${T}_sanity.log
  export SM3_TC_BWD_V=1
fi
timeout 2000 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
echo "ref rc=$?"
CMD="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/${T}_plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/${T}_launches_cfg2.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/${T}_plain_cfg2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:infonce_tc -s 6 -c 4 -o gpurun_out/${T}_ncu_cfg2 $CMD > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
kill $SMI 2>/dev/null
ls -la gpurun_out | tail -20

"""Small-shape probes on one GPU (printed as JSON lines):
  * forward POLY sweep at D=128 (the MUFU-bound width)
  * drop-in term latency at the reference's real scale (N=48..256 per rank) vs the reference op sequence on the GPU
"""
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, ".")
import skin_sm3_b200 as sm3  # noqa: E402
from oracle import ref_port  # noqa: E402


def ev_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


if len(sys.argv) > 1 and sys.argv[1] == "poly":
    n, d, T = 16384, 128, 0.1
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.randn(2 * n, d, generator=g, device="cuda")
    z, _ = sm3.core.normalize_pair(p, None, torch.bfloat16)
    ms = ev_time(lambda: sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC))
    print(json.dumps({"probe": "fwd_poly_d128", "poly": os.environ.get("SM3_TC_POLY", "0"), "ms": ms,
                      "tflops": 2 * (2 * n) ** 2 * d / ms / 1e9}))
    sys.exit(0)

for p in ("0", "2", "3"):
    r = subprocess.run([sys.executable, __file__, "poly"], env=dict(os.environ, SM3_TC_POLY=p), capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:])

crit = torch.nn.CrossEntropyLoss()
for n in (48, 64, 256, 1024):
    d, T = 128, 0.1
    g = torch.Generator().manual_seed(n)
    p1 = torch.randn(n, d, generator=g).cuda(); p2 = torch.randn(n, d, generator=g).cuda()

    def ours(prec):
        a = p1.detach().requires_grad_(True); b = p2.detach().requires_grad_(True)
        lo, la = sm3.cal_logits(a, b, T, precision=prec)
        crit(lo, la).backward()

    def ref():
        ref_port.port_infonce_step(p1, p2, T)

    out = {"probe": "term_latency", "n_pairs": n, "d": d}
    for name, fn in (("ours_fp32_us", lambda: ours("fp32")), ("ours_bf16_us", lambda: ours("bf16")), ("reference_port_gpu_us", ref)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        out[name] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
    print(json.dumps(out))

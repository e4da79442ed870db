"""Run every tcgen05 probe variant in its own process (a trapped kernel poisons the CUDA context) and print a table.
Usage (GPU box): python tools/probe_umma.py"""
import subprocess
import sys

CODE = r'''
import sys, torch
sys.path.insert(0, ".")
from skin_sm3_b200 import _lib
variant, n, k = map(int, sys.argv[1:4])
g = torch.Generator().manual_seed(1)
a = torch.randn(128, k, generator=g).bfloat16()
b = torch.randn(*((k, n) if variant & 2 else (n, k)), generator=g).bfloat16()
ac, bc = a.cuda(), b.cuda()
c = torch.full((128, n), float("nan"), device="cuda")
rc = _lib.lib().sm3_debug_umma_probe(ac.data_ptr(), bc.data_ptr(), c.data_ptr(), n, k, variant, torch.cuda.current_stream().cuda_stream)
assert rc == 0, _lib.last_error()
torch.cuda.synchronize()
ref = a.double() @ (b.double() if variant & 2 else b.double().T)
d = (c.cpu().double() - ref).abs()
print("RESULT variant=%d n=%d k=%d maxerr=%.3e refmax=%.2f nan=%d" % (variant, n, k, d.nan_to_num(1e9).max().item(), ref.abs().max().item(), int(torch.isnan(c).sum())))
'''
for variant in (0, 1, 2, 3):
    for n, k in ((128, 64), (64, 128), (256, 256)):
        try:
            r = subprocess.run([sys.executable, "-c", CODE, str(variant), str(n), str(k)], capture_output=True,
                               text=True, timeout=120)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(out[0] if out else f"FAIL variant={variant} n={n} k={k} rc={r.returncode} :: {r.stdout[-300:]} {r.stderr[-600:]}")
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT variant={variant} n={n} k={k}")

"""Pipeline timelines (clock64 of CTA (0,0)) of the D <= 128 tensor-core kernels from the -DSM3_TRACE build:
    make -C skin_sm3_b200/csrc TRACE=1 && SM3_LIB_PATH=skin_sm3_b200/lib/libsm3_b200_trace.so python tools/trace_cfg2.py [n d]
backward kinds: 0 MMA: H ready | 1 MMA: dZ issued | 2 MMA: next S issued | 3 SM: S ready | 4 SM: S in regs | 5 SM: H computed |
                6 SM: H stored + signalled | 7: CTA phases (it = 0 start, 1 first MMA, 2 last MMA issued, 3 dZ complete, 4 end)
forward  kinds: 0 MMA: tile's B landed | 1/2 MMA: S(rb0)/S(rb1) issued | 3/5 SM: S(rb) ready | 4/6 SM: S(rb) in regs, stage freed"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import skin_sm3_b200 as sm3  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
T = 0.1
lib = sm3.lib()
lib.sm3_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]


def read():
    buf = (ctypes.c_longlong * (512 * 12))()
    assert lib.sm3_debug_read_trace(buf, 512 * 12) == 0
    return np.array(buf, dtype=np.int64).reshape(512, 12)


g = torch.Generator(device="cuda").manual_seed(0)
p = torch.randn(2 * n, d, generator=g, device="cuda")
z, _ = sm3.core.normalize_pair(p, None, torch.bfloat16)
pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
gp = torch.full((2 * n,), -1e-5, device="cuda"); gl = torch.full((2 * n,), 1e-5, device="cuda")

for ver in os.environ.get("TRACE_BWD_VERSIONS", "3,2").split(","):
    os.environ["SM3_TC_BWD_V"] = ver
    sm3.reload_env()
    for _ in range(3):
        ws, k = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC)
    torch.cuda.synchronize()
    t = read()
    t0 = t[0, 7]
    ph = t[:5, 7] - t0
    nt = int((t[:, 0] > 0).sum())
    print(f"\n=== backward v{ver}  n={n} d={d}: {nt} tiles in CTA (0,0); splits={k}")
    if ver != "1":
        print(f"CTA phases (cycles from CTA start): first MMA {ph[1]}, last MMA issued {ph[2]}, dZ complete {ph[3]}, CTA end {ph[4]}")
    names = ["Hrdy", "dZiss", "Siss", "Srdy", "Sreg", "Hcmp", "Hsto"]
    base = t0 if ver != "1" else t[0][t[0] > 0].min()
    print("tile " + " ".join(f"{x:>7s}" for x in names) + "   | d(Hrdy)  ld   cmp    st  Srdy->Hsto  Hsto->dZiss")
    for it in list(range(0, min(nt, 14))) + list(range(max(14, nt - 4), nt)):
        r = t[it] - base
        prev = t[it - 1] - base if it else r
        print(f"{it:4d} " + " ".join(f"{int(x):7d}" for x in r[:7]) +
              f"   | {int(r[0] - prev[0]):6d} {int(r[4] - r[3]):5d} {int(r[5] - r[4]):5d} {int(r[6] - r[5]):5d} {int(r[6] - r[3]):7d} {int(r[1] - r[6]):7d}")
    if nt > 24:
        print("steady-state cycles per tile (MMA thread, tiles 8..nt-8):", (t[nt - 8, 0] - t[8, 0]) / (nt - 16.0))

for _ in range(3):
    sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
torch.cuda.synchronize()
t = read()
t0 = t[0, 7]
ph = t[:5, 7] - t0
nt = int((t[:, 0] > 0).sum())
print(f"\n=== forward (fwd2 if 256-row CTAs)  n={n} d={d}: {nt} tiles in CTA (0,0)")
print(f"CTA phases: first MMA {ph[1]}, last MMA issued {ph[2]}, CTA end {ph[4]}")
print("tile   Bland S0free  S0iss S1free  S1iss | S0rdy S0reg S1rdy S1reg | period  ld0   exp0(S0reg->S1rdy-ish)  ld1 | (symmetric kernel) exp1  colsum")
for it in list(range(0, min(nt, 16))):
    r = t[it] - t0
    prev = t[it - 1] - t0 if it else r
    print(f"{it:4d} {int(r[0]):7d} {int(r[8]):6d} {int(r[1]):6d} {int(r[9]):6d} {int(r[2]):6d} | {int(r[3]):6d} {int(r[4]):6d} {int(r[5]):6d} {int(r[6]):6d} | "
          f"{int(r[3] - prev[3]):6d} {int(r[4] - r[3]):5d} {int(r[5] - r[4]):6d} {int(r[6] - r[5]):5d}" +
          (f" | {int(r[10] - r[6]):6d} {int(r[11] - r[10]):6d}" if t[it, 10] > 0 else ""))
nf = int((t[:, 3] > 0).sum())
print("tiles traced in the forward CTA:", nf)

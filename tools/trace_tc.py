"""Print the per-tile pipeline timeline of infonce_tc_bwd_kernel (CTA 0,0) from a -DSM3_TRACE build.
    make -C skin_sm3_b200/csrc TRACE=1 && SM3_LIB_PATH=skin_sm3_b200/lib/libsm3_b200_trace.so python tools/trace_tc.py
kinds: 0 MMA: H ready | 1 MMA: dZ issued | 2 MMA: next S issued | 3 SM: S ready | 4 SM: S in regs | 5 SM: H computed | 6 SM: H stored+signalled"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import skin_sm3_b200 as sm3  # noqa: E402

n, d, T = 32768, 256, 0.1
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.randn(2 * n, d, generator=g, device="cuda")
z, _ = sm3.core.normalize_pair(p, None, torch.bfloat16)
pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
gp = torch.full((2 * n,), -1e-5, device="cuda"); gl = torch.full((2 * n,), 1e-5, device="cuda")
for _ in range(2):
    ws, k = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC)
torch.cuda.synchronize()
lib = sm3.lib()
MODE = sys.argv[1] if len(sys.argv) > 1 else "bwd"
if MODE == "fwd":
    for _ in range(2):
        sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * (512 * 8))()
lib.sm3_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.sm3_debug_read_trace(buf, 512 * 8) == 0
t = np.array(buf, dtype=np.int64).reshape(512, 8)
t0 = t[0][t[0] > 0].min()
if MODE == "fwd":
    buf = (ctypes.c_longlong * (512 * 8))()
    lib.sm3_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert lib.sm3_debug_read_trace(buf, 512 * 8) == 0
    t = np.array(buf, dtype=np.int64).reshape(512, 8)
    t0 = t[4, 0]
    print("fwd: tile  MMAstart issue | Srdy  Sreg  done | period  softmax(prev done->done)")
    for it in range(4, 30):
        r = t[it] - t0; prev = t[it - 1] - t0
        print(f"{it:4d} {int(r[0]):7d} {int(r[2]-r[0]):6d} | {int(r[3]):7d} {int(r[4]):7d} {int(r[5]):7d} | "
              f"{int(r[0] - prev[0]):6d} {int(r[5] - prev[5]):6d}")
    print("steady-state cycles per tile (MMA thread):", (t[200, 2] - t[40, 2]) / 160.0)
    sys.exit(0)
names = ["Hrdy", "dZiss", "Siss", "Srdy", "Sreg", "Hcmp", "Hsto"]
print("tile " + " ".join(f"{x:>7s}" for x in names) + "   | d(Hrdy) softmax(Srdy->Hsto) ld cmp st")
for it in range(4, 40):
    r = t[it] - t0
    prev = t[it - 1] - t0
    print(f"{it:4d} " + " ".join(f"{int(x):7d}" for x in r[:7]) +
          f"   | {int(r[0] - prev[0]):6d} {int(r[6] - r[3]):6d} {int(r[4] - r[3]):5d} {int(r[5] - r[4]):5d} {int(r[6] - r[5]):5d}")
per = (t[200, 0] - t[40, 0]) / 160.0
print("steady-state cycles per tile (MMA thread):", per)

"""Symmetric forward vs full forward, row by row (debug aid): python tools/sym_debug.py [n] [d]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skin_sm3_b200 as sm3  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = torch.Generator().manual_seed(5)
z, _ = sm3.core.normalize_pair(torch.randn(2 * n, d, generator=g).cuda(), None, torch.bfloat16)
res = {}
for mode in ("0", "2"):
    os.environ["SM3_TC_FWD_SYM"] = mode
    sm3.reload_env()
    pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, 0.1, sm3.ALGO_TC)
    torch.cuda.synchronize()
    res[mode] = (pos.clone().cpu(), nsum.clone().cpu())
rel = ((res["2"][1] - res["0"][1]).abs() / res["0"][1].abs())
print(f"n={n} d={d} M={2*n} T={2*n//128}: neg_sum max rel err {rel.max().item():.3e}; pos max abs err "
      f"{(res['2'][0] - res['0'][0]).abs().max().item():.3e}")
for t in range(2 * n // 128):
    r = rel[t * 128:(t + 1) * 128]
    print(f"  row tile {t:3d}: max {r.max().item():.3e} mean {r.mean().item():.3e}  ratio sym/full "
          f"{(res['2'][1][t*128:(t+1)*128] / res['0'][1][t*128:(t+1)*128]).mean().item():.4f}")
    if t > 12:
        break

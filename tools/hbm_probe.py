"""Run each HBM-bound kernel (K1 fwd/bwd, K4, K5) a few times at a bandwidth-sized shape (for ncu captures)."""
import sys

import torch

sys.path.insert(0, ".")
import skin_sm3_b200 as sm3  # noqa: E402

M, D = 1 << 21, 256
p = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
dz = torch.randn(M, D, device="cuda", dtype=torch.float32)
B = 1 << 22
x = torch.randn(B, 24, device="cuda", dtype=torch.bfloat16, requires_grad=True)
y = torch.stack([torch.randint(0, c, (B,), device="cuda") for c in sm3.NUM_CLASSES], 1)
t = torch.nn.functional.one_hot(y[:, 0], 24).to(torch.bfloat16)
for _ in range(3):
    z, inv = sm3.core.normalize_pair(p, None, torch.bfloat16)
    sm3.core.normalize_bwd(dz, 1, 1.0, z, inv, M, 0, torch.bfloat16)
    sm3.multihead_ce(x, y)
    sm3.bce_with_logits(x, t)
torch.cuda.synchronize()
print("ok")

"""tcgen05 bring-up probe: one 128 x n x k MMA chain per operand source / layout used by the real kernels
(A from smem or TMEM; B K-major or MN-major, SWIZZLE_128B).  bf16 in, fp32 accumulate => exact vs fp64 within
accumulation order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n,k", [(128, 64), (64, 128), (256, 256), (128, 256)])
def test_umma_probe(variant, n, k):
    from skin_sm3_b200 import _lib
    g = torch.Generator().manual_seed(variant * 100 + n + k)
    a = torch.randn(128, k, generator=g).bfloat16()
    b_shape = (k, n) if variant & 2 else (n, k)
    b = torch.randn(*b_shape, generator=g).bfloat16()
    ac, bc = a.cuda(), b.cuda()
    c = torch.full((128, n), float("nan"), device="cuda")
    rc = _lib.lib().sm3_debug_umma_probe(ac.data_ptr(), bc.data_ptr(), c.data_ptr(), n, k, variant,
                                         torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "probe")
    torch.cuda.synchronize()
    ref = a.double() @ (b.double() if variant & 2 else b.double().T)
    err = (c.cpu().double() - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"variant {variant} n={n} k={k}: max err {err}"

"""Parity tests proper: the CUDA path (through the C ABI / autograd drop-ins) against the golden fixtures
produced by the real reference and against the CPU oracle on seeded inputs.  Run with ``-m gpu`` on a B200.

Tolerances (BASELINE.json north_star): fp32 loss 1e-5 relative, gradients 1e-4 relative;
bf16 loss and gradients 2e-2 relative.  "relative" for a gradient tensor = max|delta| / max|reference|.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import sm3_oracle as O  # noqa: E402  (checker only)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL = ["infonce_n4_d8_T05", "infonce_n64_d128_T01", "infonce_n48_d128_T01_corr"]
BIG = ["infonce_n200_d64_T02_corr", "infonce_n512_d256_T01"]
LOSS_TOL = {"fp32": 1e-5, "bf16": 2e-2}
GRAD_TOL = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.fixture(scope="module")
def sm3():
    import skin_sm3_b200 as m
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert m.lib().sm3_device_supported() == 1, "not an sm_100 device"
    return m


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def cuda(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def relerr(a, ref):
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


def run_term(sm3, p1, p2, T, precision):
    a = p1.clone().requires_grad_(True)
    b = p2.clone().requires_grad_(True)
    logits, labels = sm3.cal_logits(a, b, T, precision=precision)
    assert logits.shape == (2 * p1.shape[0], 2) and logits.dtype == torch.float32
    assert labels.dtype == torch.long and int(labels.abs().sum()) == 0
    loss = F.cross_entropy(logits, labels)          # the script's own nn.CrossEntropyLoss()
    loss.backward()
    return loss.item(), a.grad.float().cpu().numpy(), b.grad.float().cpu().numpy(), logits.detach()


# ---------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["0", "1", "2"])     # SM3_K1_FWD_VARIANT: per-block | persistent | + evict-first
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(7, 8), (33, 5), (64, 128), (130, 256), (9, 512), (5, 1030), (70001, 256)])
def test_l2norm_forward_backward(sm3, monkeypatch, dtype, shape, variant):
    if variant != "0" and not (shape[1] % 8 == 0 and shape[1] <= 256):
        pytest.skip("the persistent forward only exists for D % 8 == 0, D <= 256")
    monkeypatch.setenv("SM3_K1_FWD_VARIANT", variant)
    g = torch.Generator().manual_seed(1)
    p = torch.randn(*shape, generator=g).to(dtype)
    p[1] = 0                                            # eps-clamp row
    pc = p.cuda().requires_grad_(True)
    z = sm3.l2_normalize(pc)
    w = torch.randn(*shape, generator=g).cuda()
    (z.float() * w).sum().backward()
    ref_p = p.double().requires_grad_(True)
    ref = F.normalize(ref_p, dim=1)
    (ref * w.cpu().double()).sum().backward()
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    assert relerr(z.detach().float().cpu(), ref.detach()) < tol
    gref = ref_p.grad.numpy()
    keep = np.ones(shape[0], bool); keep[1] = False      # clamped row: gradient is w / eps (1e12 scale)
    assert relerr(pc.grad.float().cpu().numpy()[keep], gref[keep]) < (2e-5 if dtype == torch.float32 else 2e-2)
    if dtype != torch.float16:                           # w / 1e-12 overflows fp16, in the reference too
        assert relerr(pc.grad.float().cpu().numpy()[1], gref[1]) < 1e-2
    assert not z[1].any()


# ---------------------------------------------------------------------------------------------------
# K2 / K3 against the reference-generated goldens
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", FULL + BIG)
def test_infonce_matches_reference_golden(sm3, name, precision):
    g = load(name)
    T, n = float(g["temperature"]), int(g["n"])
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    p1, p2 = cuda(g["p1"], dt), cuda(g["p2"], dt)
    loss, d1, d2, logits = run_term(sm3, p1, p2, T, precision)
    if precision == "fp32":
        ref_loss = float(g["loss_f64"])
        ref1 = g["dp1_f64"] if "dp1_f64" in g else None
    else:   # oracle evaluated on the bf16-rounded inputs the kernel actually saw
        ref_loss, r1, r2 = O.infonce_closed_form(p1.float().cpu().numpy(), p2.float().cpu().numpy(), T)
        ref1 = r1
    assert abs(loss - ref_loss) <= LOSS_TOL[precision] * max(abs(ref_loss), 1e-3), (loss, ref_loss)
    if precision == "fp32":
        if ref1 is not None:
            assert relerr(d1, g["dp1_f64"]) < GRAD_TOL[precision]
            assert relerr(d2, g["dp2_f64"]) < GRAD_TOL[precision]
            col0 = g["logits_f64"][:, 0]
        else:
            r = g["grad_rows"]
            assert relerr(d1[r], g["dp1_rows_f64"]) < GRAD_TOL[precision]
            assert relerr(d2[r], g["dp2_rows_f64"]) < GRAD_TOL[precision]
            assert relerr(d1.sum(0), g["dp1_sum_f64"]) < 1e-3
            col0 = g["logits_col0_f64"]
        # column 0 of our logits == the reference's positives column
        assert relerr(logits[:, 0].cpu().numpy(), col0) < 1e-5
    else:
        assert relerr(d1, ref1) < GRAD_TOL[precision]
        assert relerr(d2, r2) < GRAD_TOL[precision]


def test_infonce_edge_cases_fp32(sm3):
    g = load("infonce_edge")
    for case in g["case_names"]:
        for T in (0.1, 0.5):
            k = f"{case}_T{T}"
            loss, d1, d2, _ = run_term(sm3, cuda(g[k + "_p1"]), cuda(g[k + "_p2"]), T, "fp32")
            ref = float(g[k + "_loss"])
            assert abs(loss - ref) <= 1e-5 * max(abs(ref), 1.0), (k, loss, ref)
            assert np.isfinite(d1).all() and np.isfinite(d2).all(), k
            assert relerr(d1, g[k + "_dp1"]) < 1e-4, k
            assert relerr(d2, g[k + "_dp2"]) < 1e-4, k


@pytest.mark.parametrize("bwd_v", ["1", "2", "3", "4"])
@pytest.mark.parametrize("fwd_bm", ["128", "256"])
@pytest.mark.parametrize("n,d,T", [(1024, 128, 0.1), (1000, 64, 0.5), (333, 192, 0.2), (1536, 256, 0.1),
                                   (64, 256, 0.1), (129, 128, 0.07)])
def test_tensor_core_path_vs_oracle(sm3, monkeypatch, n, d, T, fwd_bm, bwd_v):
    """tcgen05 kernels (bf16 rows) vs the fp64 closed form on the same bf16-rounded inputs; ragged tile edges.
    fwd_bm selects the 128-row or the 256-row-per-CTA forward kernel (the latter is the default at cfg4 scale);
    bwd_v the backward form (1: softmax warps split the tile's columns, 2: tile-alternating groups + a_j in smem,
    3: the same with 128-column tiles where D <= 128, 4: 128-column tiles with the groups splitting each tile's columns)."""
    monkeypatch.setenv("SM3_TC_FWD_BM", fwd_bm)
    monkeypatch.setenv("SM3_TC_BWD_V", bwd_v)
    sm3.reload_env()
    g = torch.Generator().manual_seed(n + d)
    p1 = torch.randn(n, d, generator=g).bfloat16()
    p2 = (p1.float() + 0.5 * torch.randn(n, d, generator=g)).bfloat16()
    loss, d1, d2, logits = run_term(sm3, p1.cuda(), p2.cuda(), T, "bf16")
    ref_loss, r1, r2 = O.infonce_closed_form(p1.float().numpy(), p2.float().numpy(), T)
    assert abs(loss - ref_loss) <= 2e-2 * max(abs(ref_loss), 1e-3), (loss, ref_loss)
    assert relerr(d1, r1) < 2e-2 and relerr(d2, r2) < 2e-2
    # retrieval contract: the positive is found exactly where the reference finds it (argmax agreement)
    z, _ = O.normalize(np.concatenate([p1.float().numpy(), p2.float().numpy()]))
    pos_ref, lse_ref = O.infonce_stats(z, n, T)
    assert relerr(logits[:, 0].cpu().numpy(), pos_ref) < 2e-2
    assert np.abs(logits[:, 1].cpu().numpy() - lse_ref).max() < 2e-2
    # the one-call fused step (a_j materialised by the loss kernel, no prep launch) through the same kernels
    a, b = p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(True)
    l2 = sm3.fused_infonce(a, b, T, precision="bf16")
    l2.backward()
    assert abs(l2.item() - ref_loss) <= 2e-2 * max(abs(ref_loss), 1e-3)
    assert relerr(a.grad.float().cpu().numpy(), r1) < 2e-2 and relerr(b.grad.float().cpu().numpy(), r2) < 2e-2
    monkeypatch.undo()
    sm3.reload_env()


@pytest.mark.parametrize("nq", ["4", "2"])
@pytest.mark.parametrize("n,d", [(128, 128), (384, 64), (640, 256), (1152, 128), (2304, 128), (2432, 192), (4096, 256)])
def test_symmetric_forward_matches_full_forward(sm3, monkeypatch, n, d, nq):
    """The symmetric forward (upper-triangular tiles only, column sums standing in for the transposed tiles) gives the
    same row statistics as the full-matrix kernel and as the fp64 closed form, at sizes that cover an odd and an even
    number of row pairs, one tile per CTA (forced below the size threshold) and several, and every embedding width.
    Also through the fused step (loss kernel folding the triangular workspace) and deterministic run to run."""
    T = 0.1
    g = torch.Generator().manual_seed(29 + n)
    p1 = torch.randn(n, d, generator=g)
    p2 = p1 + 0.5 * torch.randn(n, d, generator=g)
    z, _ = sm3.core.normalize_pair(p1.cuda(), p2.cuda(), torch.bfloat16)
    out = {}
    try:
        for mode in ("0", "2"):
            monkeypatch.setenv("SM3_TC_FWD_SYM", mode)
            monkeypatch.setenv("SM3_TC_SYM_NQ", nq)              # softmax warps per lane quadrant of the symmetric kernel
            sm3.reload_env()
            pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
            a, b = p1.bfloat16().cuda().requires_grad_(True), p2.bfloat16().cuda().requires_grad_(True)
            loss = sm3.fused_infonce(a, b, T, precision="bf16")
            loss.backward()
            out[mode] = (pos.clone(), lse.clone(), nsum.clone(), loss.item(), a.grad.clone())
            if mode == "2":
                pos2, lse2, nsum2 = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
                assert torch.equal(nsum2, out[mode][2]) and torch.equal(pos2, out[mode][0])
    finally:
        monkeypatch.undo()
        sm3.reload_env()
    full, sym = out["0"], out["2"]
    assert torch.equal(full[0], sym[0]) or (full[0] - sym[0]).abs().max().item() < 1e-5          # positives
    assert ((full[2] - sym[2]).abs() / full[2].abs()).max().item() < 2e-5                          # neg_sum, every row
    assert (full[1] - sym[1]).abs().max().item() < 2e-5                                            # lse
    assert abs(full[3] - sym[3]) <= 2e-6 * abs(full[3])
    assert relerr(sym[4].float().cpu(), full[4].float().cpu()) < 8e-3          # bf16 gradients: one rounding step apart
    zf = z.float().cpu().numpy().astype(np.float64)
    ref_pos, ref_lse = O.infonce_stats(zf, n, T)                 # fp64 on the same bf16 rows the kernels read
    assert np.abs(sym[0].cpu().numpy() - ref_pos).max() < 1e-4
    assert np.abs(sym[1].cpu().numpy() - ref_lse).max() < 1e-4


@pytest.mark.parametrize("world,n_local,d", [(2, 256, 128), (2, 384, 64), (3, 256, 256), (4, 384, 128), (8, 256, 128), (5, 128, 192)])
def test_cross_rank_symmetric_forward_on_one_gpu(sm3, world, n_local, d):
    """Mode 4 of the multi-rank step (every rank computes world/2 of its column blocks and ships the column sums of the
    foreign ones) with this process playing all ranks on one device, production kernels throughout: each rank's K2 +
    column-sum push into per-rank statistics / flag buffers, then each rank's loss kernel.  The assembled row statistics
    must equal the full-matrix single-rank forward of the concatenated batch: odd and even world sizes, odd numbers of
    row pairs per rank, every antipodal split."""
    import ctypes as C
    from skin_sm3_b200 import _lib
    lib = _lib.lib()
    T = 0.1
    n_global = n_local * world
    g = torch.Generator().manual_seed(41 + world * 7 + n_local)
    p1 = torch.randn(n_global, d, generator=g)
    p2 = p1 + 0.5 * torch.randn(n_global, d, generator=g)
    z_all, _ = sm3.core.normalize_pair(p1.cuda(), p2.cuda(), torch.bfloat16)        # [2 n_global, d], global row order
    ref_pos, ref_lse, ref_nsum = sm3.core.stats_fwd(z_all, z_all, n_global, 0, n_global, T, sm3.ALGO_TC)
    epoch = 7
    dev = z_all.device
    stats = [torch.zeros(8 * n_global, dtype=torch.float32, device=dev) for _ in range(world)]
    flags = [torch.zeros(128, dtype=torch.int32, device=dev) for _ in range(world)]
    for f in flags:
        f[:world] = epoch                                           # channel 0: every source's rows "have landed"
    sp = (C.c_void_p * world)(*[t.data_ptr() for t in stats])
    fp = (C.c_void_p * world)(*[t.data_ptr() for t in flags])
    st = torch.cuda.current_stream().cuda_stream
    rows_of = lambda r: torch.cat([torch.arange(r * n_local, (r + 1) * n_local),          # noqa: E731
                                   n_global + torch.arange(r * n_local, (r + 1) * n_local)]).to(dev)
    ws, pos, zl = [], [], []
    for r in range(world):
        nbytes = lib.sm3_debug_mr_workspace(n_local, world, r)
        assert nbytes > 0
        ws.append(torch.empty(nbytes, dtype=torch.uint8, device=dev))
        pos.append(torch.full((2 * n_local,), float("nan"), device=dev))
        zl.append(z_all[rows_of(r)].contiguous())
        rc = lib.sm3_debug_mr_forward(zl[r].data_ptr(), z_all.data_ptr(), n_local, world, r, d, 1.0 / T, flags[r].data_ptr(),
                                      sp, fp, epoch, pos[r].data_ptr(), ws[r].data_ptr(), nbytes, st)
        assert rc == 0, _lib.last_error()
    for r in range(world):
        out = [torch.empty(2 * n_local, device=dev) for _ in range(3)]
        loss = torch.zeros((), device=dev)
        bws = torch.empty(4096, device=dev)
        rc = lib.sm3_debug_mr_fold(ws[r].data_ptr(), n_local, world, r, 1.0 / T, pos[r].data_ptr(), stats[r].data_ptr(),
                                   flags[r].data_ptr(), epoch, loss.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                   out[2].data_ptr(), bws.data_ptr(), st)
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        idx = rows_of(r)
        assert torch.equal(pos[r], ref_pos[idx]) or (pos[r] - ref_pos[idx]).abs().max().item() < 1e-5, (world, r)
        rel = ((out[0] - ref_nsum[idx]).abs() / ref_nsum[idx].abs()).max().item()
        assert rel < 2e-5, (world, n_local, r, rel)
        x = ref_lse[idx] - ref_pos[idx]
        ref_loss = torch.nn.functional.softplus(x).mean().item()
        assert abs(loss.item() - ref_loss) <= 2e-6 * abs(ref_loss), (world, r, loss.item(), ref_loss)


@pytest.mark.parametrize("bwd_v,ns", [("1", "4"), ("2", "4"), ("2", "2"), ("3", "2"), ("4", "2")])
def test_backward_forms_agree(sm3, monkeypatch, bwd_v, ns):
    """Both backward kernels and both S/H stage counts produce the same partial-gradient sums (bf16 H, fp32 accumulation:
    differences are tile-order only) on a multi-split problem with ragged edges, against the FMA kernel."""
    n, d, T = 2100, 128, 0.1
    g = torch.Generator().manual_seed(13)
    z, _ = sm3.core.normalize_pair(torch.randn(2 * n, d, generator=g).cuda(), None, torch.bfloat16)
    pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
    _, gp, gl = sm3.core.loss(pos, lse, 1.0 / (2 * n))
    ws, k = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_SIMT)
    ref = sm3.core.sum_partials(ws, k, 2 * n, d).clone()
    monkeypatch.setenv("SM3_TC_BWD_V", bwd_v)
    monkeypatch.setenv("SM3_TC_BWD_NS", ns)
    sm3.reload_env()
    try:
        ws, k = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC)
        got = sm3.core.sum_partials(ws, k, 2 * n, d).clone()
    finally:
        monkeypatch.undo()
        sm3.reload_env()
    assert relerr(got.cpu(), ref.cpu()) < 1.5e-2


def test_fp16_inputs_keep_gradient_precision_under_gradscaler(sm3):
    """--amp hands the loss fp16 projector outputs and GradScaler multiplies the upstream gradient by 65536
    (tools/backbone_train.py:98-127).  The fused steps compute the backward eagerly with upstream 1: the gradients must be
    kept in fp32 until that factor has been applied (per-element values ~1e-7 at N = 4096 are below fp16's normal range)."""
    n, d, T, scale = 4096, 128, 0.1, 65536.0
    gen = torch.Generator().manual_seed(17)
    p1 = torch.randn(n, d, generator=gen).half()
    p2 = (p1.float() + 0.5 * torch.randn(n, d, generator=gen)).half()
    ref_loss, r1, r2 = O.infonce_closed_form(p1.float().numpy(), p2.float().numpy(), T, chunk=2048)
    a, b = p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(True)
    loss = sm3.fused_infonce(a, b, T, precision="bf16")
    (loss * scale).backward()
    assert a.grad.dtype == torch.float16 and b.grad.dtype == torch.float16
    assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss)
    assert relerr(a.grad.float().cpu().numpy() / scale, r1) < 2e-2
    assert relerr(b.grad.float().cpu().numpy() / scale, r2) < 2e-2
    # the small-magnitude tail must survive too (it is what an fp16 intermediate would flush): rows' gradient norms
    gn = (a.grad.float() / scale).norm(dim=1).cpu().numpy()
    rn = np.linalg.norm(r1, axis=1)
    assert np.abs(gn - rn).max() <= 2e-2 * rn.max()
    # grouped form
    a2, b2 = p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(True)
    l2 = sm3.fused_infonce_multi([(a2, b2), (b2, a2)], T, [1.0, 0.5], precision="bf16")
    (l2 * scale).backward()
    assert a2.grad.dtype == torch.float16
    assert relerr(a2.grad.float().cpu().numpy() / scale, 1.5 * r1) < 2e-2


def test_range_limits_raise_cleanly(sm3):
    """Documented limits of the CUDA path: embedding width <= 256 and 1/T < 83.  There is no
    CPU / eager fallback by design, so both must raise a RuntimeError that names the limit -- not compute garbage."""
    p = torch.randn(32, 512, device="cuda")
    with pytest.raises(RuntimeError, match="D=512"):
        sm3.cal_logits(p, p.clone(), 0.1, precision="fp32")
    q = torch.randn(32, 128, device="cuda")
    with pytest.raises(RuntimeError, match="temperature"):
        sm3.cal_logits(q, q.clone(), 0.005, precision="fp32")
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        sm3.cal_logits(q.cpu(), q.cpu(), 0.1)


def test_stage_timing_reports_the_production_kernels(sm3):
    import ctypes as C
    n, d, T = 1024, 128, 0.1
    a = torch.randn(n, d, device="cuda").bfloat16().requires_grad_(True)
    b = torch.randn(n, d, device="cuda").bfloat16().requires_grad_(True)
    lib = sm3.lib()
    lib.sm3_stage_timing(1)
    try:
        sm3.fused_infonce(a, b, T, precision="bf16").backward()
        buf = (C.c_float * 8)()
        k = lib.sm3_stage_timing_read(buf, 8)
        names = lib.sm3_stage_timing_names().decode().split(",")
    finally:
        lib.sm3_stage_timing(0)
    assert k == 5 and names == ["normalize", "infonce_fwd", "loss", "infonce_bwd", "normalize_bwd"]
    assert all(0.0 < buf[i] < 50.0 for i in range(k))


def test_tensor_core_and_fma_kernels_agree_on_identical_rows(sm3):
    """Same bf16 rows through both CUDA kernels: differences are accumulation-order only."""
    n, d, T = 700, 128, 0.1
    g = torch.Generator().manual_seed(7)
    p = torch.randn(2 * n, d, generator=g).cuda()
    z, _ = sm3.core.normalize_pair(p, None, torch.bfloat16)
    a = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
    b = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_SIMT)
    for x, y in zip(a, b):
        assert relerr(x.cpu(), y.cpu()) < 1e-4
    gp = torch.randn(2 * n, generator=g).cuda() * 1e-3
    gl = torch.rand(2 * n, generator=g).cuda() * 1e-3
    outs = []
    for algo in (sm3.ALGO_TC, sm3.ALGO_SIMT):
        ws, k = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, b[2], gp, gl, b[2], algo)
        outs.append(sm3.core.sum_partials(ws, k, 2 * n, d).clone())
    assert relerr(outs[0].cpu(), outs[1].cpu()) < 1.5e-2     # H is rounded to bf16 on the tensor-core path


def test_row_block_sharding_reproduces_single_block(sm3):
    """W-rank emulation on one GPU: per-rank row blocks with global column order == one-block result."""
    n, d, T, W = 512, 128, 0.1, 4
    g = torch.Generator().manual_seed(11)
    p = torch.randn(2 * n, d, generator=g).cuda()
    for z_dtype, algo in ((torch.bfloat16, sm3.ALGO_TC), (torch.float32, sm3.ALGO_SIMT)):
        z, _ = sm3.core.normalize_pair(p, None, z_dtype)
        pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, algo)
        gp = torch.randn(2 * n, generator=g).cuda() * 1e-3
        gl = torch.rand(2 * n, generator=g).cuda() * 1e-3
        ws, k = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, algo)
        dz = sm3.core.sum_partials(ws, k, 2 * n, d).clone()
        nl = n // W
        for r in range(W):
            rows = torch.from_numpy(O.global_row_index(nl, r * nl, n)).cuda()
            zr = z[rows].contiguous()
            ps, ls, ns = sm3.core.stats_fwd(zr, z, nl, r * nl, n, T, algo)
            assert relerr(ps.cpu(), pos[rows].cpu()) < 1e-5 and relerr(ns.cpu(), nsum[rows].cpu()) < 1e-5
            assert relerr(ls.cpu(), lse[rows].cpu()) < 1e-5
            ws, k = sm3.core.stats_bwd(zr, z, nl, r * nl, n, T, gp[rows].contiguous(), gl[rows].contiguous(),
                                       nsum[rows].contiguous(), gp, gl, nsum, algo)
            dzr = sm3.core.sum_partials(ws, k, 2 * nl, d)
            assert relerr(dzr.cpu(), dz[rows].cpu()) < 1e-4


@pytest.mark.parametrize("n_local,d,world", [(512, 128, 4), (1024, 256, 2), (256, 64, 8)])
def test_owner_ordered_forward_matches_plain_forward(sm3, monkeypatch, n_local, d, world):
    """Mode 3 of the multi-rank step visits the column tiles owner by owner (own columns, then rank-1's, ...), every split
    taking every S-th tile of an owner's range, and waits for one flag per source rank.  Emulated on one GPU: the test
    plays the peers (complete column buffer, flags pre-set), every rank's row block must reproduce the plain kernel."""
    monkeypatch.setenv("SM3_TC_FWD_BM", "256")
    T = 0.1
    n = n_local * world
    g = torch.Generator().manual_seed(n + d)
    z, _ = sm3.core.normalize_pair(torch.randn(2 * n, d, generator=g).cuda(), None, torch.bfloat16)
    pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
    flags = torch.full((128,), 7, dtype=torch.int32, device="cuda")
    lib = sm3.lib()
    for r in range(world):
        rows = torch.from_numpy(O.global_row_index(n_local, r * n_local, n)).cuda()
        zr = z[rows].contiguous()
        out = torch.empty((3, 2 * n_local), dtype=torch.float32, device="cuda")
        ws = torch.empty(int(lib.sm3_debug_infonce_fwd_ordered_workspace(n_local, world, d)) + 256, dtype=torch.uint8, device="cuda")
        rc = lib.sm3_debug_infonce_fwd_ordered(zr.data_ptr(), z.data_ptr(), n_local, r, world, d, 1.0 / T, flags.data_ptr(), 7,
                                               out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), ws.data_ptr(),
                                               ws.numel(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, sm3._lib.last_error()
        assert relerr(out[0].cpu(), pos[rows].cpu()) < 1e-5, ("pos", r)
        assert relerr(out[2].cpu(), nsum[rows].cpu()) < 1e-5, ("neg_sum", r)       # same tiles, another summation order
        assert relerr(out[1].cpu(), lse[rows].cpu()) < 1e-5, ("lse", r)


def test_fused_scalar_loss_and_grad_scaling(sm3):
    """fused_infonce == CE(cal_logits) incl. an upstream scale (GradScaler x loss weight, backbone_train.py:101-125)."""
    g = load("infonce_n64_d128_T01")
    T = float(g["temperature"])
    p1 = cuda(g["p1"]).requires_grad_(True); p2 = cuda(g["p2"]).requires_grad_(True)
    loss = sm3.fused_infonce(p1, p2, T, precision="fp32")
    (loss * 65536.0 * 0.5).backward()
    assert abs(loss.item() - float(g["loss_f64"])) < 1e-5 * float(g["loss_f64"])
    assert relerr(p1.grad.cpu().numpy() / 32768.0, g["dp1_f64"]) < 1e-4
    assert relerr(p2.grad.cpu().numpy() / 32768.0, g["dp2_f64"]) < 1e-4


def test_fused_step_entry_matches_composed_path(sm3):
    """sm3_infonce_step (one C call) == the Python-composed sequence of the same kernels (gradients bit for bit)."""
    from skin_sm3_b200 import functional as F3
    g = load("infonce_n200_d64_T02_corr")
    T = float(g["temperature"])
    for prec, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        outs = []
        for profiled in (False, True):
            F3._PROFILE = (lambda name: None) if profiled else None
            try:
                a = cuda(g["p1"], dt).requires_grad_(True); b = cuda(g["p2"], dt).requires_grad_(True)
                loss = sm3.fused_infonce(a, b, T, precision=prec)
                loss.backward()
                outs.append((loss.item(), a.grad.clone(), b.grad.clone()))
            finally:
                F3._PROFILE = None
        # the one-call step folds the loss with a multi-CTA kernel, the composed path with a single CTA: same terms,
        # different fp32 summation order; the per-row gradients are computed identically
        assert abs(outs[0][0] - outs[1][0]) <= 1e-6 * abs(outs[1][0])
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_grouped_terms_match_separate_calls(sm3):
    """sm3_infonce_step_multi: the four style-0 terms (derm + clinic + 0.5 cross1 + 0.5 cross2) in one call ==
    four fused_infonce calls: gradients bit for bit, loss to fp32 summation order; and against the oracle."""
    n, d, T = 200, 128, 0.1
    gen = torch.Generator().manual_seed(5)
    w = [1.0, 1.0, 0.5, 0.5]
    for dt, prec, tl, tg in ((torch.bfloat16, "bf16", 2e-2, 2e-2), (torch.float32, "fp32", 1e-5, 1e-4)):
        base = [(torch.randn(n, d, generator=gen).to(dt), torch.randn(n, d, generator=gen).to(dt)) for _ in range(4)]
        ps = [(a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)) for a, b in base]
        loss = sm3.fused_infonce_multi(ps, T, w, precision=prec)
        (loss * 3.0).backward()
        qs = [(a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)) for a, b in base]
        sep = sum(wi * sm3.fused_infonce(a, b, T, precision=prec) for wi, (a, b) in zip(w, qs))
        (sep * 3.0).backward()
        assert abs(loss.item() - sep.item()) <= 1e-6 * abs(sep.item())
        ref_total = 0.0
        for wi, (a, b), (qa, qb), (ha, hb) in zip(w, ps, qs, base):
            assert torch.equal(a.grad, qa.grad) and torch.equal(b.grad, qb.grad)
            rl, r1, r2 = O.infonce_closed_form(ha.float().numpy(), hb.float().numpy(), T)
            ref_total += wi * rl
            assert relerr(a.grad.float().cpu().numpy() / (3.0 * wi), r1) < tg
        assert abs(loss.item() - ref_total) <= tl * abs(ref_total)
    with pytest.raises(ValueError):
        sm3.fused_infonce_multi([], T)
    with pytest.raises(ValueError):
        sm3.fused_infonce_multi(ps[:2], T, [1.0])


def test_cuda_graph_step_matches_eager(sm3):
    """GraphedInfoNCE (the one-call step captured into a CUDA graph) == the eager op, bit for bit, across replays with
    fresh inputs written into the static buffers."""
    n, d, T = 320, 128, 0.1
    gen = torch.Generator().manual_seed(21)
    for dt, prec in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
        gr = sm3.GraphedInfoNCE(n, d, T, dtype=dt, precision=prec, weight=0.5)
        for _ in range(3):
            a = torch.randn(n, d, generator=gen).to(dt).cuda()
            b = torch.randn(n, d, generator=gen).to(dt).cuda()
            gr.p1.copy_(a); gr.p2.copy_(b)
            loss, d1, d2 = gr.replay()
            ae, be = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
            le = sm3.fused_infonce(ae, be, T, precision=prec, weight=0.5)
            le.backward()
            assert loss.item() == le.item()
            assert torch.equal(d1, ae.grad) and torch.equal(d2, be.grad)


def test_host_buffer_entry(sm3):
    g = load("infonce_n64_d128_T01")
    T, n, d = float(g["temperature"]), int(g["n"]), int(g["d"])
    for dt, algo, tl, tg in ((torch.float32, sm3.ALGO_SIMT, 1e-5, 1e-4), (torch.bfloat16, sm3.ALGO_TC, 2e-2, 2e-2)):
        h = sm3.HostInfoNCE(n, d, dt, algo)
        p1 = torch.from_numpy(g["p1"]).to(dt).pin_memory(); p2 = torch.from_numpy(g["p2"]).to(dt).pin_memory()
        loss, d1, d2 = h(p1, p2, T)
        ref_loss, r1, r2 = O.infonce_closed_form(p1.float().numpy(), p2.float().numpy(), T)
        assert abs(loss.item() - ref_loss) < tl * ref_loss
        assert relerr(d1.float().numpy(), r1) < tg and relerr(d2.float().numpy(), r2) < tg


def test_host_pipeline_matches_synchronous_entry(sm3):
    """sm3_host_pipe_* (three streams, slots reused every `depth` submits) == sm3_infonce_host on the same batches, bit
    for bit, with several different batches in flight and more tickets than slots."""
    n, d, T = 384, 128, 0.1
    gen = torch.Generator().manual_seed(11)
    batches = [(torch.randn(n, d, generator=gen).bfloat16().pin_memory(),
                torch.randn(n, d, generator=gen).bfloat16().pin_memory()) for _ in range(7)]
    sync = sm3.HostInfoNCE(n, d, torch.bfloat16, sm3.ALGO_TC)
    want = []
    for p1, p2 in batches:
        l, d1, d2 = sync(p1, p2, T)
        want.append((l.clone(), d1.clone(), d2.clone()))
    for depth in (1, 2, 3):
        pipe = sm3.HostInfoNCEPipeline(n, d, torch.bfloat16, sm3.ALGO_TC, depth)
        tickets, got = [], []
        for p1, p2 in batches:
            tickets.append(pipe.submit(p1, p2, T))
            if len(tickets) == depth:
                got.append(tuple(t.clone() for t in pipe.wait(tickets.pop(0))))
        for t in tickets:
            got.append(tuple(x.clone() for x in pipe.wait(t)))
        pipe.close()
        assert len(got) == len(want)
        for (l, d1, d2), (rl, r1, r2) in zip(got, want):
            assert torch.equal(l, rl) and torch.equal(d1, r1) and torch.equal(d2, r2)
    # forward-only tickets and argument validation
    pipe = sm3.HostInfoNCEPipeline(n, d, torch.bfloat16, sm3.ALGO_TC, 2)
    with pytest.raises(RuntimeError, match="ticket"):
        pipe.wait(5)
    pipe.close()
    with pytest.raises(ValueError):
        sm3.HostInfoNCEPipeline(n, d, torch.bfloat16, sm3.ALGO_TC, 9)


def test_known_answers_at_full_size(sm3):
    """Size-independent properties at BASELINE config-4 scale (N=32768, D=256, M=65536 rows)."""
    n, d, T = 32768, 256, 0.1
    m = 2 * n
    # (i) all rows identical -> every similarity is 1 -> loss = log(M-1)
    p = torch.ones(n, d, device="cuda", dtype=torch.bfloat16)
    loss = sm3.fused_infonce(p, p, T, precision="bf16")
    assert abs(loss.item() - np.log(m - 1)) < 1e-3
    # (ii) z_{i+N} = z_i = e_{i mod D}: positives 1, D-fold duplicates among the negatives
    idx = torch.arange(n, device="cuda") % d
    e = torch.zeros(n, d, device="cuda", dtype=torch.bfloat16)
    e[torch.arange(n, device="cuda"), idx] = 1
    loss = sm3.fused_infonce(e, e, T, precision="bf16")
    same = m // d - 2                    # negatives identical to the row, rest orthogonal
    expect = np.log(np.exp(1 / T) * (1 + same) + (m - 2 - same)) - 1 / T
    assert abs(loss.item() - expect) < 1e-3 * expect
    # (iii) random rows at the benchmarked geometry (fwd2<4> grid (256, 4), bwd<4> grid (512, 2) on 148 SMs): loss and
    # the gradient of 640 rows -- first / last row tiles, the rows either side of the two halves' boundary, and 384 rows
    # spread over the rest -- against oracle.infonce_rowblock (fp32-GEMM mode, pinned to the reference goldens to 1e-5
    # by tests/test_oracle_golden.py).  Every row's gradient sums over ALL 65536 columns, i.e. over every column split.
    g = torch.Generator(device="cuda").manual_seed(3407)
    p1 = torch.randn(n, d, generator=g, device="cuda").bfloat16().requires_grad_(True)
    p2 = torch.randn(n, d, generator=g, device="cuda").bfloat16().requires_grad_(True)
    loss = sm3.fused_infonce(p1, p2, T, precision="bf16")       # == the call bench.py times (sm3_infonce_step)
    loss.backward()
    rng = np.random.default_rng(5)
    rows = np.unique(np.concatenate([np.arange(64), np.arange(n - 64, n + 64), np.arange(m - 64, m),
                                     rng.choice(m, 384, replace=False)]))
    ref_loss, ref_dp, ref_lse = O.infonce_rowblock(p1.detach().float().cpu().numpy(), p2.detach().float().cpu().numpy(),
                                                   T, rows, matmul_dtype=np.float32)
    assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss), (loss.item(), ref_loss)
    assert abs(loss.item() - ref_loss) <= 1e-3 * abs(ref_loss), "bf16 rows, fp32 accumulation: expected ~1e-4"
    got = torch.cat([p1.grad, p2.grad]).float()[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert relerr(got, ref_dp) < 2e-2, relerr(got, ref_dp)
    # per-row statistics of ALL 65536 rows through the drop-in path (K2 + finalize): log(e^pos + e^lse_neg) == lse
    logits, _ = sm3.cal_logits(p1.detach(), p2.detach(), T, precision="bf16")
    lse = torch.logaddexp(logits[:, 0], logits[:, 1]).cpu().numpy()
    assert np.abs(lse - ref_lse).max() < 2e-2, np.abs(lse - ref_lse).max()
    # (iv) tensor-core path == FMA path on a 256-row block of the same problem
    z, _ = sm3.core.normalize_pair(p1.detach(), p2.detach(), torch.bfloat16)
    nl = 128
    rows = torch.from_numpy(O.global_row_index(nl, 5 * nl, n)).cuda()
    zr = z[rows].contiguous()
    a = sm3.core.stats_fwd(zr, z, nl, 5 * nl, n, T, sm3.ALGO_TC)
    b = sm3.core.stats_fwd(zr, z, nl, 5 * nl, n, T, sm3.ALGO_SIMT)
    for x, y in zip(a, b):
        assert relerr(x.cpu(), y.cpu()) < 1e-4


def test_cfg2_full_parity(sm3):
    """BASELINE configs[1] at its own size (N = 4096 pairs, D = 128, bf16): loss and EVERY gradient element of the fused
    step (eager op and CUDA-graph replay, the two calls bench.py times) against the fp64 closed form."""
    n, d, T = 4096, 128, 0.1
    gen = torch.Generator().manual_seed(3407)
    p1 = torch.randn(n, d, generator=gen).bfloat16()
    p2 = (p1.float() + 0.5 * torch.randn(n, d, generator=gen)).bfloat16()
    ref_loss, r1, r2 = O.infonce_closed_form(p1.float().numpy(), p2.float().numpy(), T, chunk=2048)
    a, b = p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(True)
    loss = sm3.fused_infonce(a, b, T, precision="bf16")
    loss.backward()
    assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss), (loss.item(), ref_loss)
    assert relerr(a.grad.float().cpu().numpy(), r1) < 2e-2 and relerr(b.grad.float().cpu().numpy(), r2) < 2e-2
    gr = sm3.GraphedInfoNCE(n, d, T, dtype=torch.bfloat16, precision="bf16")
    gr.p1.copy_(p1.cuda()); gr.p2.copy_(p2.cuda())
    l2, d1, d2 = gr.replay()
    assert abs(l2.item() - ref_loss) <= 2e-2 * abs(ref_loss)
    assert relerr(d1.float().cpu().numpy(), r1) < 2e-2 and relerr(d2.float().cpu().numpy(), r2) < 2e-2
    # uncorrelated pairs (the bench's synthetic batch): same check on the loss and 512 gradient rows
    q1, q2 = torch.randn(n, d, generator=gen).bfloat16(), torch.randn(n, d, generator=gen).bfloat16()
    rows = np.arange(0, 2 * n, 16)
    ref_loss, ref_dp, _ = O.infonce_rowblock(q1.float().numpy(), q2.float().numpy(), T, rows)
    a, b = q1.cuda().requires_grad_(True), q2.cuda().requires_grad_(True)
    loss = sm3.fused_infonce(a, b, T, precision="bf16")
    loss.backward()
    assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss)
    got = torch.cat([a.grad, b.grad]).float()[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert relerr(got, ref_dp) < 2e-2


# ---------------------------------------------------------------------------------------------------
# N2: fused projector tail (last Linear + affine-free BatchNorm + F.normalize) against the real make_projector goldens
# ---------------------------------------------------------------------------------------------------
def _tail_modules(w, rm, rv, dtype):
    d, k = w.shape
    lin = torch.nn.Linear(k, d, bias=False).cuda()
    bn = torch.nn.BatchNorm1d(d, affine=False).cuda()
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(w))
        bn.running_mean.copy_(torch.from_numpy(rm)); bn.running_var.copy_(torch.from_numpy(rv))
    return lin, bn


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_projector_tail_kernels_vs_reference_golden(sm3, dtype):
    """sm3_proj_tail_gemm / _bn_l2 / _bwd1 / _bwd2 one by one: z, the running statistics and d/dh, d/dW of sum(z * G)
    against values recorded from the real layers (fp64); eval mode too."""
    import ctypes as C
    lib = sm3.lib()
    g = load("projtail")
    code = {torch.bfloat16: 2, torch.float16: 1}[dtype]
    st = torch.cuda.current_stream().cuda_stream
    for c in g["cases"]:
        h32, w32, gz = g[f"{c}_h"], g[f"{c}_w"], g[f"{c}_gz"]
        r, k = h32.shape
        d = w32.shape[0]
        h, w = cuda(h32, dtype), cuda(w32, dtype)                     # bf16-representable values: exact in fp16 too? (no:
        href, wref = h.float().cpu().numpy().astype(np.float64), w.float().cpu().numpy().astype(np.float64)   # re-read)
        f = O.projector_tail(href, wref, running=(g[f"{c}_rm0"], g[f"{c}_rv0"]))
        y = torch.empty((r, d), dtype=torch.float32, device="cuda")
        totals = torch.empty(2 * d, dtype=torch.float32, device="cuda")
        ws = torch.empty(int(lib.sm3_proj_tail_workspace_bytes(r, d)), dtype=torch.uint8, device="cuda")
        assert lib.sm3_proj_tail_gemm(h.data_ptr(), w.data_ptr(), r, k, d, code, y.data_ptr(), totals.data_ptr(), ws.data_ptr(),
                                      ws.numel(), st) == 0, sm3._lib.last_error()
        assert relerr(y.cpu().numpy(), f["y"]) < 1e-5                  # fp32 accumulation of exact 16-bit products
        assert relerr(totals[:d].cpu().numpy(), f["y"].sum(0)) < 1e-4 and relerr(totals[d:].cpu().numpy(), (f["y"] ** 2).sum(0)) < 1e-4
        rm, rv = cuda(g[f"{c}_rm0"]), cuda(g[f"{c}_rv0"])
        mean, rstd = torch.empty(d, device="cuda"), torch.empty(d, device="cuda")
        z = torch.empty((r, d), dtype=torch.bfloat16, device="cuda")
        inv = torch.empty(r, dtype=torch.float32, device="cuda")
        assert lib.sm3_proj_tail_bn_l2(y.data_ptr(), r, d, totals.data_ptr(), float(r), 1e-5, 1e-12, 1, 0.1, rm.data_ptr(),
                                       rv.data_ptr(), mean.data_ptr(), rstd.data_ptr(), z.data_ptr(), inv.data_ptr(), st) == 0
        assert relerr(z.float().cpu().numpy(), f["z"]) < 8e-3          # bf16 output
        assert relerr(inv.cpu().numpy(), f["inv"]) < 1e-4
        assert relerr(rm.cpu().numpy(), f["running_mean"]) < 1e-5 and relerr(rv.cpu().numpy(), f["running_var"]) < 1e-4
        # backward through the same kernels: upstream dz = G in one "partial slab"
        dz = cuda(gz)
        dyhat = torch.empty((r, d), dtype=torch.float32, device="cuda")
        totals2 = torch.empty(2 * d, dtype=torch.float32, device="cuda")
        assert lib.sm3_proj_tail_bwd1(dz.data_ptr(), 1, r * d, z.data_ptr(), inv.data_ptr(), 1e-12, y.data_ptr(), mean.data_ptr(),
                                      rstd.data_ptr(), r, d, dyhat.data_ptr(), totals2.data_ptr(), ws.data_ptr(), ws.numel(), st) == 0
        dy = torch.empty((r, d), dtype=torch.float32, device="cuda")
        assert lib.sm3_proj_tail_bwd2(dyhat.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), totals2.data_ptr(),
                                      float(r), 1, r, d, dy.data_ptr(), 0, st) == 0
        dh_ref, dw_ref, dy_ref = O.projector_tail_bwd(href, wref, f, gz)
        assert relerr(dy.cpu().numpy(), dy_ref) < 2e-2                 # z is bf16
        assert relerr((dy.double() @ w.double()).cpu().numpy(), dh_ref) < 2e-2
        assert relerr((dy.double().t() @ h.double()).cpu().numpy(), dw_ref) < 2e-2
        # eval mode: running statistics
        rm1, rv1 = cuda(g[f"{c}_rm1"]), cuda(g[f"{c}_rv1"])
        assert lib.sm3_proj_tail_bn_l2(y.data_ptr(), r, d, None, float(r), 1e-5, 1e-12, 0, 0.0, rm1.data_ptr(), rv1.data_ptr(),
                                       mean.data_ptr(), rstd.data_ptr(), z.data_ptr(), inv.data_ptr(), st) == 0
        fe = O.projector_tail(href, wref, running=(g[f"{c}_rm1"], g[f"{c}_rv1"]), training=False)
        assert relerr(z.float().cpu().numpy(), fe["z"]) < 8e-3


def test_tail_cal_logits_matches_reference_term(sm3):
    """Two real projector tails -> _cal_logits -> CE (golden): loss and gradients w.r.t. both inputs and both weights
    through skin_sm3_b200.tail_cal_logits + the script's own criterion, and the one-group form against cal_logits."""
    g = load("projtail")
    n, T = int(g["term_n"]), float(g["term_T"])
    tails, leaves = [], []
    for i in (1, 2):
        lin, bn = _tail_modules(g[f"term_w{i}"], np.zeros(g[f"term_w{i}"].shape[0]), np.ones(g[f"term_w{i}"].shape[0]), torch.bfloat16)
        h = cuda(g[f"term_f{i}"], torch.bfloat16).requires_grad_(True)
        assert sm3.tail_supported(h, lin, bn)
        tails.append(sm3.TailSpec(h, lin.weight.to(torch.bfloat16), bn))
        leaves.append((h, lin, bn))
    logits, labels = sm3.tail_cal_logits(tails, T)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    ref = float(g["term_loss"])
    assert abs(loss.item() - ref) <= 2e-2 * abs(ref), (loss.item(), ref)
    for i, (h, lin, bn) in enumerate(leaves, 1):
        assert relerr(h.grad.float().cpu().numpy(), g[f"term_df{i}"]) < 2e-2, i
        assert relerr(lin.weight.grad.float().cpu().numpy(), g[f"term_dw{i}"]) < 2e-2, i
        assert int(bn.num_batches_tracked) == 1
    # one group of 2N rows == the stock layers + cal_logits on the same rows (SimCLR.forward's joint BatchNorm)
    c = "b"
    lin, bn = _tail_modules(g[f"{c}_w"], g[f"{c}_rm0"], g[f"{c}_rv0"], torch.bfloat16)
    lin2, bn2 = _tail_modules(g[f"{c}_w"], g[f"{c}_rm0"], g[f"{c}_rv0"], torch.bfloat16)
    h = cuda(g[f"{c}_h"], torch.bfloat16).requires_grad_(True)
    h2 = cuda(g[f"{c}_h"], torch.bfloat16).requires_grad_(True)
    la, _ = sm3.tail_cal_logits([sm3.TailSpec(h, lin.weight.to(torch.bfloat16), bn)], T)
    p = bn2(F.linear(h2.float(), lin2.weight))
    half = p.shape[0] // 2
    lb, _ = sm3.cal_logits(p[:half], p[half:], T, precision="bf16")
    F.cross_entropy(la, torch.zeros(len(la), dtype=torch.long, device="cuda")).backward()
    F.cross_entropy(lb, torch.zeros(len(lb), dtype=torch.long, device="cuda")).backward()
    assert relerr(la[:, 1].detach().cpu(), lb[:, 1].detach().cpu()) < 2e-2
    assert relerr(h.grad.float().cpu(), h2.grad.float().cpu()) < 3e-2
    assert relerr(bn.running_var.cpu(), bn2.running_var.cpu()) < 1e-3


# ---------------------------------------------------------------------------------------------------
# heads
# ---------------------------------------------------------------------------------------------------
def test_multihead_ce_matches_reference_loops(sm3):
    g = load("heads")
    nc = [int(c) for c in g["num_classes"]]
    x = cuda(g["eval_logits"]).requires_grad_(True)
    loss = sm3.multihead_ce(list(torch.split(x, nc, dim=1)), cuda(g["eval_labels"], torch.long),
                            weights=list(g["eval_weights"]))
    loss.backward()
    assert abs(loss.item() - float(g["eval_loss"])) < 1e-5 * float(g["eval_loss"])
    assert relerr(x.grad.cpu().numpy(), g["eval_grad"]) < 1e-4
    x = cuda(g["dc_logits"]).requires_grad_(True)
    loss = sm3.multihead_ce(x, cuda(g["dc_labels"], torch.long), temperature=float(g["dc_temperature"]),
                            ignore_index=-100, class_counts=nc)
    loss.backward()
    assert abs(loss.item() - float(g["dc_loss"])) < 1e-5 * float(g["dc_loss"])
    assert relerr(x.grad.cpu().numpy(), g["dc_grad"]) < 1e-4


@pytest.mark.parametrize("variant", ["0", "1"])          # SM3_CE_VARIANT: synchronous slab | cp.async double buffer
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B", [1, 64, 257, 4096, 100003, 700001])
def test_multihead_ce_vs_oracle(sm3, monkeypatch, dtype, B, variant):
    monkeypatch.setenv("SM3_CE_VARIANT", variant)
    rng = np.random.default_rng(B)
    nc = list(sm3.NUM_CLASSES)
    x = torch.from_numpy(rng.normal(size=(B, 24)).astype(np.float32) * 2).to(dtype)
    y = np.stack([rng.integers(0, c, B) for c in nc], axis=1)
    y[rng.random((B, 8)) < 0.1] = -100
    if (y[:, 0] == -100).all():
        y[0, :] = 0
    w = [1, 2, 0.5, 1, 1, 3, 1, 0.25]
    xc = x.cuda().requires_grad_(True)
    loss = sm3.multihead_ce(xc, torch.from_numpy(y).cuda(), weights=w, temperature=0.5, ignore_index=-100)
    loss.backward()
    ref, gref = O.multihead_ce(x.float().numpy(), y, w, 2.0, -100, nc)
    if np.isnan(ref):
        assert np.isnan(loss.item())
        return
    assert abs(loss.item() - ref) < 2e-5 * abs(ref)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    if dtype == torch.float16 and B > 200000:
        # d loss / d logit ~ 1 / (8 B) ~ 1e-7 sits at the fp16 subnormal step (6e-8): the stored gradient is quantised
        # to a few percent whatever computes it (torch's own fp16 CE backward included); use a GradScaler-like factor
        xs = x.cuda().requires_grad_(True)
        (sm3.multihead_ce(xs, torch.from_numpy(y).cuda(), weights=w, temperature=0.5, ignore_index=-100) * 1024.0).backward()
        assert relerr(xs.grad.float().cpu().numpy() / 1024.0, gref) < tol
        return
    assert relerr(xc.grad.float().cpu().numpy(), gref) < tol


@pytest.mark.parametrize("variant", ["0", "1"])          # SM3_BCE_VARIANT: one log per element | per 8 elements
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 24), (37, 24), (512, 24), (1001, 5), (65536, 24), (300007, 24)])
def test_bce_with_logits_vs_torch_oracle(sm3, monkeypatch, dtype, shape, variant):
    monkeypatch.setenv("SM3_BCE_VARIANT", variant)
    rng = np.random.default_rng(shape[0])
    x = torch.from_numpy(rng.normal(size=shape).astype(np.float32) * 3).to(dtype)
    t = torch.from_numpy((rng.random(shape) < 0.3).astype(np.float32))
    pw = torch.from_numpy(rng.random(shape[1]).astype(np.float32) * 3 + 0.5)
    for pos_weight in (None, pw):
        xc = x.cuda().requires_grad_(True)
        loss = sm3.bce_with_logits(xc, t.cuda(), None if pos_weight is None else pos_weight.cuda())
        loss.backward()
        ref, gref = O.bce_with_logits(x.float().numpy(), t.numpy(), None if pos_weight is None else pw.numpy())
        assert abs(loss.item() - ref) < 1e-5 * abs(ref)
        assert relerr(xc.grad.float().cpu().numpy(), gref) < (1e-4 if dtype == torch.float32 else 1e-2)


# ---------------------------------------------------------------------------------------------------
# drop-in module against the real reference wrappers (tiny encoder golden)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls_name", ["SimCLRSkinV3", "SimCLRSkinV32"])
def test_dropin_model_matches_reference(sm3, cls_name):
    from skin_sm3_b200 import dropin
    dropin.install()
    from src.models import resnet as my_resnet
    from src.models import simclr as mine
    from tiny_encoder import tiny
    my_resnet.__dict__["tiny"] = tiny
    g = load("model_tiny")
    N, PD, T = int(g["N"]), int(g["proj_dim"]), float(g["temperature"])
    model = getattr(mine, cls_name)("tiny", weights=None, proj_dim=PD, temperature=T)
    keys = [str(k) for k in g[f"{cls_name}/sd_keys"]]
    assert list(model.state_dict().keys()) == keys
    sd = {k: torch.from_numpy(g[f"{cls_name}/sd/{k}"]) for k in keys}
    imgs = [cuda(g["imgs"][i]) for i in range(4)]
    crit = torch.nn.CrossEntropyLoss().cuda()          # tools/backbone_train.py:531
    for style, wts in ((0, (0.5, 0.5)), (1, (0.5, 0.5)), (2, (0.25,) * 4)):
        model.load_state_dict(sd)
        model = model.float().cuda().train()
        model.zero_grad(set_to_none=True)
        out = model([imgs[0], imgs[1]], [imgs[2], imgs[3]], style)
        assert len(out) == 3 and len(out[2]) == len(wts)
        cross = sum(w * crit(*out[2][k]) for k, w in enumerate(wts))
        derm, clinic = crit(*out[0]), crit(*out[1])
        loss = derm + clinic + cross                     # tools/backbone_train.py:98-121
        loss.backward()
        pre = f"{cls_name}/style{style}/"
        assert abs(derm.item() - float(g[pre + "derm_loss"])) < 1e-4 * abs(float(g[pre + "derm_loss"]))
        assert abs(clinic.item() - float(g[pre + "clinic_loss"])) < 1e-4 * abs(float(g[pre + "clinic_loss"]))
        got = np.array([crit(*o).item() for o in out[2]])
        assert np.abs(got - g[pre + "cross_losses"]).max() < 1e-4 * np.abs(g[pre + "cross_losses"]).max()
        assert abs(loss.item() - float(g[pre + "loss"])) < 1e-4 * abs(float(g[pre + "loss"]))
        for k, p in model.named_parameters():
            ref = g[pre + "grad/" + k]
            if ref.size == 0:
                assert p.grad is None or not p.grad.any()
                continue
            assert relerr(p.grad.cpu().numpy(), ref) < 2e-3, (style, k)


# ---------------------------------------------------------------------------------------------------
# N1 retrieval
# ---------------------------------------------------------------------------------------------------
def test_knn_topk_matches_reference_evaluator(sm3):
    g = load("knn")
    q, bank = cuda(g["query"]), cuda(g["bank"])
    k = int(g["k"])
    vals, idx = sm3.sim_topk(q, bank, k)
    assert (idx.cpu().numpy() == g["topk_idx"]).all()            # exact index agreement with torch.topk of the reference
    assert relerr(vals.cpu().numpy(), g["topk_val"]) < 1e-6
    pred = sm3.knn_predict(q, bank, cuda(g["bank_labels"], torch.long), int(g["num_classes"]), k,
                           float(g["temperature"]))
    assert (pred.cpu().numpy() == g["pred_labels"]).all()


@pytest.mark.parametrize("form", ["1", "2"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("bq,nb,d,k", [(5, 7, 12, 3), (64, 1500, 128, 200), (257, 4096, 256, 50), (33, 1025, 64, 256)])
def test_sim_topk_vs_oracle(sm3, monkeypatch, dtype, bq, nb, d, k, form):
    monkeypatch.setenv("SM3_TOPK_TILED", form)      # 1: tiled + threshold filter, 2: materialised + radix select
    rng = np.random.default_rng(bq + nb)
    q = torch.from_numpy(rng.normal(size=(bq, d)).astype(np.float32)).to(dtype)
    bank = torch.from_numpy(rng.normal(size=(nb, d)).astype(np.float32)).to(dtype)
    bank[3] = bank[1]                                             # exact tie -> lower index first
    vals, idx = sm3.sim_topk(q.cuda(), bank.cuda(), k)
    rv, ri = O.knn_topk(q.float().numpy(), bank.float().numpy(), k)
    v, i = vals.cpu().numpy(), idx.cpu().numpy()
    assert relerr(v, rv) < 1e-5
    sim = q.double().numpy() @ bank.double().numpy().T
    for r in range(bq):
        if (i[r] == ri[r]).all():
            continue
        # any disagreement must be a near-tie in fp32 accumulation order
        bad = np.nonzero(i[r] != ri[r])[0]
        assert np.abs(sim[r, i[r][bad]] - sim[r, ri[r][bad]]).max() < 1e-4 * np.abs(sim[r]).max(), (r, bad)
    both = np.sort(i, axis=1)
    assert (both[:, 1:] != both[:, :-1]).all()                    # no duplicates


@pytest.mark.parametrize("bq,nb,d,k,self_off", [(512, 16384, 128, 200, -1), (1000, 3000, 96, 5, 0), (40, 70000, 64, 17, -1)])
def test_sim_topk_forms_agree(sm3, monkeypatch, bq, nb, d, k, self_off):
    """The tiled, threshold-filtered search (bank split across CTAs + merge kernel) and the single-pass kernel return the
    same top-k set: same 64-bit keys, so any difference can only come from fp32 accumulation order of near-ties."""
    g = torch.Generator().manual_seed(bq + k)
    q = torch.randn(bq, d, generator=g).cuda()
    bank = torch.randn(nb, d, generator=g).cuda()
    bank[7] = bank[3]                                                        # exact ties: lower index first in every form
    monkeypatch.setenv("SM3_TOPK_TILED", "1")
    v1, i1 = sm3.sim_topk(q, bank, k, exclude_self_offset=self_off)
    monkeypatch.setenv("SM3_TOPK_TILED", "2")
    v2, i2 = sm3.sim_topk(q, bank, k, exclude_self_offset=self_off)
    monkeypatch.setenv("SM3_TOPK_TILED", "0")
    v0, i0 = sm3.sim_topk(q, bank, k, exclude_self_offset=self_off)
    monkeypatch.undo()
    assert (i2 == i0).float().mean().item() > 0.999 and relerr(v2.cpu().numpy(), v0.cpu().numpy()) < 1e-5
    assert (v2[:, :-1] >= v2[:, 1:]).all()
    s2 = torch.sort(i2, dim=1).values
    assert (s2[:, 1:] != s2[:, :-1]).all() and (i2 >= 0).all() and (i2 < nb).all()
    assert (v1[:, :-1] >= v1[:, 1:]).all()                                   # sorted descending
    assert (i1 >= 0).all() and (i1 < nb).all()
    if self_off >= 0:
        assert (i1 != (torch.arange(bq, device=i1.device) + self_off)[:, None]).all()
    agree = (i1 == i0).float().mean().item()
    assert agree > 0.999, agree
    assert relerr(v1.cpu().numpy(), v0.cpu().numpy()) < 1e-5
    s1 = torch.sort(i1, dim=1).values
    assert (s1[:, 1:] != s1[:, :-1]).all()                                   # no duplicates
    # against the library formulation on the same GPU
    sim = q @ bank.T
    if self_off >= 0:
        sim[torch.arange(bq), torch.arange(bq) + self_off] = -float("inf")
    rv, ri = sim.topk(k, dim=1)
    assert (ri == i1).float().mean().item() > 0.995
    assert relerr(v1.cpu().numpy(), rv.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("form", ["0", "1", "2"])
def test_sim_topk_many_exact_ties_take_the_lowest_indices(sm3, monkeypatch, form):
    """A bank of identical rows: every similarity ties, so the result must be indices 0 .. k-1 (1 .. k with the query's own
    row excluded) -- the ordered tie scan of the radix-select form, the 64-bit keys of the other two."""
    monkeypatch.setenv("SM3_TOPK_TILED", form)
    row = torch.randn(1, 64, generator=torch.Generator().manual_seed(2))
    bank = row.repeat(3000, 1).cuda()
    q = row.repeat(9, 1).cuda()
    v, i = sm3.sim_topk(q, bank, 17)
    assert (i.cpu() == torch.arange(17)[None, :]).all()
    v, i = sm3.sim_topk(bank[:9].contiguous(), bank, 17, exclude_self_offset=0)
    exp = torch.stack([torch.tensor([j for j in range(18) if j != r][:17]) for r in range(9)])
    assert (i.cpu() == exp).all()


def test_in_batch_retrieval_top1_is_the_positive(sm3):
    """golden: top-5 non-self neighbours of every row from the reference similarity matrix (argmax agreement)."""
    g = load("infonce_n48_d128_T01_corr")
    p = np.concatenate([g["p1"], g["p2"]])
    z, _ = sm3.core.normalize_pair(cuda(p), None, torch.float32)
    vals, idx = sm3.sim_topk(z, z, 5, exclude_self_offset=0)
    assert (idx.cpu().numpy() == g["top5_idx"]).all()
    assert relerr(vals.cpu().numpy(), g["top5_val"]) < 1e-5
    n = int(g["n"])
    assert (idx[:, 0].cpu().numpy() == (np.arange(2 * n) + n) % (2 * n)).all()   # correlated pairs: positive is top-1


# ---------------------------------------------------------------------------------------------------
# N3: fused prototype heads against the reference's multi-label Model (golden from tools/mlc_train.py:58-89, 255-261)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_proto_heads_match_reference_model(sm3, dtype):
    g = load("mlchead")
    counts = [5, 3, 2, 3, 3, 3, 3, 2]
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    for tag in g["cases"]:
        l2, T = bool(g[f"{tag}_l2"]), float(g[f"{tag}_T"])
        if dtype != torch.float32 and g[f"{tag}_sa_in"].shape[2] % 256:
            continue                                              # 16-bit rows need D % 256 == 0
        sa = cuda(g[f"{tag}_sa_in"], dtype).requires_grad_(True)
        w = cuda(g[f"{tag}_w"])
        ws = [t.clone().requires_grad_(True) for t in torch.split(w, counts, dim=0)]
        assert sm3.proto_heads_supported(sa, 24)
        sa_out, logits = sm3.proto_heads(sa, ws, l2)
        assert logits.shape == (sa.shape[1], 24) and logits.dtype == torch.float32 and sa_out.shape == sa.shape
        ref_sa = g[f"{tag}_sa_out"] if dtype == torch.float32 else O.proto_heads(sa.detach().float().cpu().numpy(), g[f"{tag}_w"], counts, l2)[0]
        ref_logits = g[f"{tag}_logits"] if dtype == torch.float32 else O.proto_heads(sa.detach().float().cpu().numpy(), g[f"{tag}_w"], counts, l2)[1]
        assert relerr(sa_out.detach().float().cpu().numpy(), ref_sa) < tol, tag
        assert relerr(logits.detach().cpu().numpy(), ref_logits) < tol, tag
        # DeepCluster loss through the fused 8-head CE (the [B, 24] layout needs no torch.cat) and backward through both
        loss = sm3.multihead_ce(logits, cuda(g[f"{tag}_targets"], torch.long), temperature=T, ignore_index=-100,
                                class_counts=counts)
        loss.backward()
        if dtype == torch.float32:
            assert abs(loss.item() - float(g[f"{tag}_loss"])) < 2e-5, tag
            assert relerr(sa.grad.cpu().numpy(), g[f"{tag}_d_sa_in"]) < 1e-4, tag
            assert relerr(torch.cat([t.grad for t in ws]).cpu().numpy(), g[f"{tag}_dw"]) < 1e-4, tag
        else:
            _, dl = O.multihead_ce(ref_logits, g[f"{tag}_targets"], None, 1.0 / T, ignore_index=-100)
            dx, dw = O.proto_heads_bwd(sa.detach().float().cpu().numpy(), g[f"{tag}_w"], counts, l2, dl)
            assert relerr(sa.grad.float().cpu().numpy(), dx) < 3e-2, tag
            assert relerr(torch.cat([t.grad for t in ws]).cpu().numpy(), dw) < 3e-2, tag
        # a gradient arriving at the returned features (not the reference's use, but legal) is added in
        sa2 = cuda(g[f"{tag}_sa_in"], dtype).requires_grad_(True)
        so2, lg2 = sm3.proto_heads(sa2, [t.detach() for t in ws], l2)
        (so2.float().sum() + lg2.sum()).backward()
        xs = cuda(g[f"{tag}_sa_in"], dtype).float().requires_grad_(True)
        zz = torch.nn.functional.normalize(xs, dim=-1) if l2 else xs
        ref = zz.sum() + sum((zz[h % zz.shape[0]] @ t.detach().t()).sum() for h, t in enumerate(ws))
        ref.backward()
        assert relerr(sa2.grad.float().cpu().numpy(), xs.grad.cpu().numpy()) < (1e-4 if dtype == torch.float32 else 3e-2), tag


# ---------------------------------------------------------------------------------------------------
# N4: DeepCluster memory-bank k-means against the reference's cluster_memory (golden)
# ---------------------------------------------------------------------------------------------------
def test_cluster_memory_matches_reference(sm3):
    import types
    g = load("kmeans")
    for c in g["cases"]:
        emb, index, init = g[f"{c}_emb"], g[f"{c}_index"], g[f"{c}_init_idx"]
        k, d = len(init), emb.shape[1]
        a, cent = sm3.spherical_kmeans(cuda(emb), torch.from_numpy(init).cuda())
        ra, rcent = O.spherical_kmeans(emb, init)
        assert (a.cpu().numpy() == ra).all(), c
        assert np.abs(cent.cpu().numpy() - g[f"{c}_centroids"]).max() < 1e-5, c
        # the drop-in: same signature and side effect as tools/mlc_train.py::cluster_memory, same RNG consumption
        proto = torch.nn.Linear(d, k, bias=False).cuda()
        torch.manual_seed(int(g[f"{c}_seed"]))
        out = sm3.cluster_memory(types.SimpleNamespace(world_size=1, rank=0), proto, k, torch.from_numpy(index).cuda(),
                                 cuda(emb))
        assert (out.cpu().numpy() == g[f"{c}_assign"]).all(), c
        assert np.abs(proto.weight.detach().cpu().numpy() - g[f"{c}_centroids"]).max() < 1e-5, c
    # k = 1 retrieval (the E step) against the oracle on its own
    q = torch.randn(300, 96, generator=torch.Generator().manual_seed(3)).cuda()
    b = torch.randn(7, 96, generator=torch.Generator().manual_seed(4)).cuda()
    v, i = sm3.sim_topk(q, b, 1)
    rv, ri = O.knn_topk(q.cpu().numpy(), b.cpu().numpy(), 1)
    assert (i.cpu().numpy() == ri).all() and relerr(v.cpu().numpy(), rv) < 1e-5

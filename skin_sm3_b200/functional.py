"""torch.autograd.Function drop-ins for the SM3 contrastive hot path, backed by libsm3_b200.so.

Reference being replaced (paths relative to the reference checkout):
  * ``l2_normalize``            F.normalize(x, dim=1)                      src/models/simclr.py:62,138,294
  * ``cal_logits``              SimCLRSkinV3._cal_logits after projectors  src/models/simclr.py:293-322
                                (= SimCLR.forward :61-88, SimCLRSkinV2._cal_logits :137-166)
  * ``fused_infonce``           the above + nn.CrossEntropyLoss + backward tools/backbone_train.py:101-125,531
  * ``multihead_ce``            the 8-head CE loops                        tools/mlc_eval.py:159-162,
                                                                           tools/mlc_train.py:255-261
  * ``bce_with_logits``         north_star's multi-hot head loss (no reference counterpart)
  * ``fused_infonce_multi``     the four style-0 terms of one step in one call  tools/backbone_train.py:101-121
  * ``sim_topk`` / ``knn_predict``  KNNOnlineEvaluator.predict                 src/models/evaluator.py:43-83
  * ``cluster_memory`` / ``spherical_kmeans``  DeepCluster memory-bank k-means tools/mlc_train.py:116-189
  * ``tail_cal_logits``         projector tail Linear + BatchNorm1d(affine=False) + normalise + the above
                                                                           src/models/simclr.py:25-26,61-62,293-294
  * ``proto_heads`` / ``mlc_model_forward``  normalise + the eight prototype Linears of Model.forward
                                                                           tools/mlc_train.py:58-89
  * ``HostInfoNCE`` / ``HostInfoNCEPipeline`` / ``GraphedInfoNCE``  host-buffer, pipelined and CUDA-graph front ends of
                                the fused step (new capability; bench.py's ``e2e`` and small-shape numbers)

``cal_logits`` returns *sufficient-statistics logits* ``[M, 2] = [pos/T, log sum_neg exp(s/T)]`` with
all-zero labels: the training script's own ``nn.CrossEntropyLoss()(logits, labels)`` then yields exactly
the reference loss, and autograd hands (g_pos, g_lse) back to the fused backward kernel.  The ``[M, M-1]``
logits matrix of the reference never exists.

Everything here requires CUDA tensors and the built extension; there is no CPU / eager fallback.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ALGO_AUTO, ALGO_SIMT, ALGO_TC, check, dtype_code, lib, ptr, require_cuda, stream_ptr

NUM_CLASSES = (5, 3, 2, 3, 3, 3, 3, 2)   # DIAG + seven-point checklist (tools/mlc_eval.py:63)
_EPS = 1e-12                             # F.normalize default


# bench.py installs a callable(name) here that records a CUDA event on the current stream after each stage
_PROFILE = None

# Scratch of the one-call fused steps, reused across calls: keyed by (kind, device, stream, shape...).  All users of
# one entry enqueue on the same stream, so consecutive steps are ordered and may share it (nothing in the scratch is
# read after the call's last kernel: loss and gradients live in their own tensors).  Saves an allocation and a ctypes
# size query per step on the launch-bound shapes.
_SCRATCH = {}


def reload_env() -> None:
    """Re-read the SM3_TC_* tuning knobs (sweeps / tests that change them inside one process): drops the library's cached
    values and the scratch cache, whose sizes depend on the kernel variant."""
    lib().sm3_debug_reload_env()
    _SCRATCH.clear()


_PLAN_KNOBS = ("SM3_TC_FWD_BM", "SM3_TC_BWD_V", "SM3_TC_BWD_NS", "SM3_TC_FWD_SPLITS", "SM3_TC_BWD_SPLITS",
               "SM3_TC_FWD_SYM")


def _scratch(kind: str, dev: torch.device, key: tuple, nbytes_fn) -> torch.Tensor:
    # the launch plans (hence the sizes) depend on the tuning knobs that are read from the environment at every call
    env = os.environ
    k = (kind, dev.index, torch.cuda.current_stream(dev).cuda_stream) + key + tuple(env.get(n) for n in _PLAN_KNOBS)
    buf = _SCRATCH.get(k)
    if buf is None:
        if len(_SCRATCH) > 64:
            _SCRATCH.clear()
        buf = torch.empty(max(int(nbytes_fn()), 256), dtype=torch.uint8, device=dev)
        _SCRATCH[k] = buf
    return buf


def _mark(name: str) -> None:
    if _PROFILE is not None:
        _PROFILE(name)


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _autocast_on() -> bool:
    try:
        return torch.is_autocast_enabled("cuda")
    except TypeError:  # older signature
        return torch.is_autocast_enabled()


def pick_precision(p: torch.Tensor, precision: str = "auto"):
    """-> (z dtype, algo).  'bf16': bf16 rows on the tcgen05 kernels (D % 64 == 0, D <= 256; other widths use
    the FMA kernels on bf16 rows).  'fp32': fp32 rows on the FMA kernels (reference-exact fp32 parity).
    'auto': fp32 inputs outside autocast -> 'fp32', everything else -> 'bf16'."""
    if precision == "auto":
        precision = "fp32" if (p.dtype == torch.float32 and not _autocast_on()) else "bf16"
    if precision == "fp32":
        return torch.float32, ALGO_SIMT
    if precision != "bf16":
        raise ValueError(f"precision must be 'auto', 'bf16' or 'fp32', got {precision!r}")
    d = p.shape[-1]
    return torch.bfloat16, (ALGO_TC if (d % 64 == 0 and 64 <= d <= 256) else ALGO_SIMT)


# ======================================================================================================
# raw (non-autograd) wrappers around the C ABI
# ======================================================================================================
class core:
    """Thin typed wrappers; every method launches on the current stream of the tensors' device."""

    @staticmethod
    def normalize_pair(p1: torch.Tensor, p2: Optional[torch.Tensor], z_dtype: torch.dtype, eps: float = _EPS):
        dev = require_cuda(p1, p2)
        p1 = _contig(p1)
        n1, d = p1.shape
        n2 = 0
        if p2 is not None:
            p2 = _contig(p2)
            if p2.shape[1] != d or p2.dtype != p1.dtype:
                raise ValueError("p1 / p2 must agree in width and dtype")
            n2 = p2.shape[0]
        z = torch.empty((n1 + n2, d), dtype=z_dtype, device=dev)
        inv = torch.empty(n1 + n2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().sm3_l2norm_fwd(ptr(p1), n1, ptr(p2), n2, d, dtype_code(p1), ptr(z), dtype_code(z), ptr(inv),
                                       eps, stream_ptr()), "sm3_l2norm_fwd")
        return z, inv

    @staticmethod
    def normalize_bwd(dz_partials: torch.Tensor, n_partials: int, scale: float, z: torch.Tensor, inv: torch.Tensor,
                      n1: int, n2: int, out_dtype: torch.dtype, eps: float = _EPS):
        dev = require_cuda(dz_partials, z, inv)
        d = z.shape[1]
        dp1 = torch.empty((n1, d), dtype=out_dtype, device=dev)
        dp2 = torch.empty((n2, d), dtype=out_dtype, device=dev) if n2 else None
        with torch.cuda.device(dev):
            check(lib().sm3_l2norm_bwd(ptr(dz_partials), n_partials, scale, ptr(z), dtype_code(z), ptr(inv), eps,
                                       ptr(dp1), n1, ptr(dp2), n2, d, dtype_code(dp1), stream_ptr()),
                  "sm3_l2norm_bwd")
        return dp1, dp2

    @staticmethod
    def workspace(n_local: int, n_global: int, d: int, z: torch.Tensor, algo: int, backward: bool) -> torch.Tensor:
        nbytes = lib().sm3_infonce_workspace_bytes(n_local, n_global, d, dtype_code(z), algo, int(backward))
        return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=z.device)

    @staticmethod
    def stats_fwd(z_rows: torch.Tensor, z_cols: torch.Tensor, n_local: int, pair_offset: int, n_global: int,
                  temperature: float, algo: int = ALGO_AUTO):
        dev = require_cuda(z_rows, z_cols)
        d = z_rows.shape[1]
        m = 2 * n_local
        assert z_rows.shape[0] == m and z_cols.shape[0] == 2 * n_global and z_cols.shape[1] == d
        assert z_rows.is_contiguous() and z_cols.is_contiguous() and z_rows.dtype == z_cols.dtype
        out = torch.empty((3, m), dtype=torch.float32, device=dev)   # pos, lse_neg, neg_sum
        ws = core.workspace(n_local, n_global, d, z_rows, algo, False)
        with torch.cuda.device(dev):
            check(lib().sm3_infonce_fwd(ptr(z_rows), ptr(z_cols), n_local, pair_offset, n_global, d,
                                        dtype_code(z_rows), 1.0 / temperature, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                        ptr(ws), ws.numel(), algo, stream_ptr()), "sm3_infonce_fwd")
        return out[0], out[1], out[2]

    @staticmethod
    def stats_bwd(z_rows, z_cols, n_local, pair_offset, n_global, temperature, g_pos_r, g_lse_r, nsum_r,
                  g_pos_c, g_lse_c, nsum_c, algo: int = ALGO_AUTO):
        """-> (fp32 partial-gradient workspace, n_partials); feed to normalize_bwd / sum_partials."""
        dev = require_cuda(z_rows, z_cols)
        d = z_rows.shape[1]
        for t in (g_pos_r, g_lse_r, nsum_r, g_pos_c, g_lse_c, nsum_c):
            assert t.dtype == torch.float32 and t.is_contiguous()
        ws = core.workspace(n_local, n_global, d, z_rows, algo, True)
        with torch.cuda.device(dev):
            npart = check(lib().sm3_infonce_bwd(ptr(z_rows), ptr(z_cols), n_local, pair_offset, n_global, d,
                                                dtype_code(z_rows), 1.0 / temperature, ptr(g_pos_r), ptr(g_lse_r),
                                                ptr(nsum_r), ptr(g_pos_c), ptr(g_lse_c), ptr(nsum_c), ptr(ws),
                                                ws.numel(), algo, stream_ptr()), "sm3_infonce_bwd")
        return ws, npart

    @staticmethod
    def stats_bwd_packed(z_rows, z_cols, n_local, pair_offset, n_global, temperature, g_pos_r, g_lse_r, nsum_r,
                         stats_cols, algo: int = ALGO_AUTO):
        """As stats_bwd with the column statistics packed as float4 rows [2*n_global, 4] (peer-exchange layout)."""
        dev = require_cuda(z_rows, z_cols, stats_cols)
        d = z_rows.shape[1]
        assert stats_cols.dtype == torch.float32 and stats_cols.shape == (2 * n_global, 4) and stats_cols.is_contiguous()
        ws = core.workspace(n_local, n_global, d, z_rows, algo, True)
        with torch.cuda.device(dev):
            npart = check(lib().sm3_infonce_bwd_packed(ptr(z_rows), ptr(z_cols), n_local, pair_offset, n_global, d,
                                                       dtype_code(z_rows), 1.0 / temperature, ptr(g_pos_r), ptr(g_lse_r),
                                                       ptr(nsum_r), ptr(stats_cols), ptr(ws), ws.numel(), algo,
                                                       stream_ptr()), "sm3_infonce_bwd_packed")
        return ws, npart

    @staticmethod
    def stats_fwd_remote(z_rows, z_cols, n_local, pair_offset, n_global, temperature, nsum_local, pos_local):
        """K2 over the column tiles owned by other ranks; adds nsum_local.  -> (lse_neg, neg_sum) totals."""
        dev = require_cuda(z_rows, z_cols)
        d = z_rows.shape[1]
        out = torch.empty((2, 2 * n_local), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nbytes = lib().sm3_infonce_remote_workspace_bytes(n_local, n_global, d, 0)
            ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
            check(lib().sm3_infonce_fwd_remote(ptr(z_rows), ptr(z_cols), n_local, pair_offset, n_global, d,
                                               dtype_code(z_rows), 1.0 / temperature, ptr(nsum_local), ptr(pos_local),
                                               ptr(out[0]), ptr(out[1]), ptr(ws), ws.numel(), stream_ptr()),
                  "sm3_infonce_fwd_remote")
        return out[0], out[1]

    @staticmethod
    def stats_bwd_split(z, z_cols, n_local, pair_offset, n_global, temperature, g_pos, g_lse, nsum, stats_cols,
                        between=None):
        """Row-local backward as local block + remote blocks into ONE slab buffer.  `between()` is called after the
        local block has been enqueued (the caller waits for the peers' statistics there).  -> (ws, n_partials)."""
        dev = require_cuda(z, z_cols)
        d = z.shape[1]
        m = 2 * n_local
        with torch.cuda.device(dev):
            b_l = lib().sm3_infonce_workspace_bytes(n_local, n_local, d, dtype_code(z), ALGO_TC, 1)
            b_r = lib().sm3_infonce_remote_workspace_bytes(n_local, n_global, d, 1)
            ws = torch.empty(int(b_l) + int(b_r) + 1024, dtype=torch.uint8, device=dev)
            np_l = check(lib().sm3_infonce_bwd(ptr(z), ptr(z), n_local, 0, n_local, d, dtype_code(z), 1.0 / temperature,
                                               ptr(g_pos), ptr(g_lse), ptr(nsum), ptr(g_pos), ptr(g_lse), ptr(nsum),
                                               ptr(ws), int(b_l), ALGO_TC, stream_ptr()), "sm3_infonce_bwd(local)")
            if between is not None:
                between()
            off = np_l * m * d * 4
            np_r = check(lib().sm3_infonce_bwd_remote_packed(ptr(z), ptr(z_cols), n_local, pair_offset, n_global, d,
                                                             dtype_code(z), 1.0 / temperature, ptr(g_pos), ptr(g_lse),
                                                             ptr(nsum), ptr(stats_cols), ws.data_ptr() + off,
                                                             ws.numel() - off, stream_ptr()),
                         "sm3_infonce_bwd_remote_packed")
        return ws, np_l + np_r

    @staticmethod
    def sum_partials(ws: torch.Tensor, n_partials: int, m: int, d: int) -> torch.Tensor:
        v = ws[: n_partials * m * d * 4].view(torch.float32).view(n_partials, m, d)
        return v[0] if n_partials == 1 else v.sum(0)

    @staticmethod
    def scale_grads(dps, g: torch.Tensor, out_dtypes):
        """out_t = dp_t * g for all tensors in ONE launch (``sm3_scale_grads``); the cast to the inputs' dtype (fp32 ->
        fp16 for --amp inputs) happens in the same kernel, after the multiplication."""
        import ctypes as C
        dps = [_contig(d) for d in dps]
        g32 = g if (g.dtype == torch.float32 and g.dim() == 0) else g.reshape(()).float()
        same = len(set(out_dtypes)) == 1 and len({d.numel() for d in dps}) == 1 and len({d.dtype for d in dps}) == 1
        if not same or len(dps) > 8:
            return tuple((d * g32.to(d.dtype)).to(o) for d, o in zip(dps, out_dtypes))
        outs = [torch.empty(d.shape, dtype=out_dtypes[0], device=d.device) for d in dps]
        k = len(dps)
        with torch.cuda.device(dps[0].device):
            check(lib().sm3_scale_grads((C.c_void_p * k)(*[d.data_ptr() for d in dps]),
                                        (C.c_void_p * k)(*[o.data_ptr() for o in outs]), k, dps[0].numel(),
                                        dtype_code(dps[0]), _lib._DTYPES[out_dtypes[0]], ptr(g32), stream_ptr()),
                  "sm3_scale_grads")
        return tuple(outs)

    @staticmethod
    def loss(pos, lse_neg, scale: float, out: Optional[torch.Tensor] = None, accumulate: bool = False,
             want_grads: bool = True):
        dev = require_cuda(pos, lse_neg)
        m = pos.numel()
        if out is None:
            out = torch.empty((), dtype=torch.float32, device=dev)
        g = torch.empty((2, m), dtype=torch.float32, device=dev) if want_grads else None
        with torch.cuda.device(dev):
            check(lib().sm3_infonce_loss(ptr(pos), ptr(lse_neg), m, scale, ptr(out), int(accumulate),
                                         ptr(g[0]) if want_grads else None, ptr(g[1]) if want_grads else None,
                                         stream_ptr()), "sm3_infonce_loss")
        return out, (g[0] if want_grads else None), (g[1] if want_grads else None)


# ======================================================================================================
# cross-rank plumbing (row-block sharding): global order = [all first views ; all second views]
# ======================================================================================================
def gather_global_order(t_local: torch.Tensor, group=None) -> torch.Tensor:
    """t_local: [2*n_local, ...] = [first halves ; second halves] of this rank.  Returns [2*n_global, ...]
    with rank r's first halves at [r*n_local, (r+1)*n_local) and its second halves n_global further on.
    Equal n_local on every rank (DistributedSampler drops/pads, reference src/utils/misc.py:400)."""
    w = dist.get_world_size(group)
    m = t_local.shape[0]
    n_local = m // 2
    t_local = _contig(t_local)
    # ONE collective (rank-major [W, 2, n_local, ...]) + one local permute copy to [2, W, n_local, ...]:
    # cheaper than two NCCL launches on the latency-bound sizes this path sees.
    stage = torch.empty((w, 2, n_local) + tuple(t_local.shape[1:]), dtype=t_local.dtype, device=t_local.device)
    dist.all_gather_into_tensor(stage.view((w * m,) + tuple(t_local.shape[1:])), t_local, group=group)
    return stage.transpose(0, 1).reshape((2 * n_local * w,) + tuple(t_local.shape[1:]))


def _grad_safe(p: torch.Tensor) -> torch.Tensor:
    """fp16 inputs (the reference's --amp dtype) are handed to the fused step as fp32: the step computes the backward
    eagerly with upstream gradient 1, and d loss / d p ~ 1 / (2N) per element would land in fp16's subnormal / flush range
    before ``backward()`` multiplies by the GradScaler factor.  fp32 storage keeps it exact; the cast back to fp16 happens
    after that multiplication (as _MultiHeadCE / _BCEWithLogits do).  bf16 / fp32 inputs pass through untouched."""
    return p.float() if p.dtype == torch.float16 else p


def _scale_grads(dps, g: torch.Tensor, out_dtypes):
    """dp * upstream for the eagerly computed gradients (see core.scale_grads)."""
    return core.scale_grads(dps, g, out_dtypes)


def _group_info(group):
    if group is None or not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


# ======================================================================================================
# autograd Functions
# ======================================================================================================
class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, eps, out_dtype):
        z, inv = core.normalize_pair(p, None, out_dtype or p.dtype, eps)
        ctx.save_for_backward(z, inv)
        ctx.eps, ctx.p_dtype = eps, p.dtype
        return z

    @staticmethod
    def backward(ctx, dz):
        z, inv = ctx.saved_tensors
        dz32 = _contig(dz.float())
        dp, _ = core.normalize_bwd(dz32, 1, 1.0, z, inv, z.shape[0], 0, ctx.p_dtype, ctx.eps)
        return dp, None, None


def l2_normalize(p: torch.Tensor, eps: float = _EPS, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Row-wise ``F.normalize(p, dim=1, eps)`` (2-D input)."""
    if p.dim() != 2:
        raise ValueError("l2_normalize expects a 2-D tensor [rows, D]")
    return _L2Normalize.apply(p, eps, out_dtype)


_LOGITS_SLOTS = 8        # peer slots of the logits form: up to 4 terms between a forward and its backward (style 0 / 2)


def _logits_peer_buffers(w: int, z_dtype, algo: int, n_global: int, d: int, device, group):
    """PeerBuffers for ``cal_logits(group=...)`` (NVLink peer-memory exchange instead of NCCL all-gathers), or None.
    SM3_LOGITS_COMM = auto | peer | nccl (auto: peer when symmetric memory can be set up)."""
    comm = os.environ.get("SM3_LOGITS_COMM", "auto")
    if w == 1 or comm == "nccl" or z_dtype != torch.bfloat16 or algo == ALGO_SIMT:
        if comm == "peer" and w > 1:
            raise RuntimeError("SM3_LOGITS_COMM=peer needs the bf16 tensor-core path")
        return None
    from . import peer
    try:
        return peer.get_peer_buffers(group, n_global, d, device, depth=_LOGITS_SLOTS)
    except peer.PeerUnavailable:
        if comm == "peer":
            raise
        return None


class _InfoNCELogits(torch.autograd.Function):
    """normalise -> exchange of the rows -> K2 statistics;  backward = exchange of 3 floats/row -> K3 -> normalise-bwd.
    The exchange is the NVLink peer-memory scatter + cross-rank barrier (symmetric memory) where it can be set up, NCCL
    all-gathers otherwise.  The peer slot of a term stays untouched until its backward has run: `_LOGITS_SLOTS` slots
    rotate, enough for the four terms SimCLRSkinV3 keeps in flight between forward and backward."""

    @staticmethod
    def forward(ctx, p1, p2, temperature, z_dtype, algo, group):
        w, rank = _group_info(group)
        n_local = p1.shape[0]
        z, inv = core.normalize_pair(p1, p2, z_dtype)
        n_global = n_local * w
        pbuf = _logits_peer_buffers(w, z_dtype, algo, n_global, p1.shape[1], p1.device, group)
        slot = -1
        if w == 1:
            z_cols = z
        elif pbuf is not None:
            slot = pbuf.next_slot()
            z_cols = pbuf.scatter_z(slot, z, n_local)          # NVLink stores into every rank's z_cols[slot] + barrier
        else:
            z_cols = gather_global_order(z, group)
        pos, lse, nsum = core.stats_fwd(z, z_cols, n_local, rank * n_local, n_global, temperature, algo)
        ctx.save_for_backward(z, inv, z_cols if pbuf is None else z.new_empty(0), nsum)
        ctx.meta = (temperature, algo, group, w, rank, n_local, p1.dtype, pbuf, slot, pbuf.step if pbuf is not None else 0)
        return torch.stack((pos, lse), dim=1)

    @staticmethod
    def backward(ctx, g):
        z, inv, z_cols, nsum = ctx.saved_tensors
        temperature, algo, group, w, rank, n_local, p_dtype, pbuf, slot, issued = ctx.meta
        g = g.float()
        g_pos, g_lse = _contig(g[:, 0]), _contig(g[:, 1])
        if pbuf is not None:
            if pbuf.step - issued >= pbuf.DEPTH:              # the slot has been handed to a later forward meanwhile
                raise RuntimeError("cal_logits: more than %d terms between a forward and its backward" % pbuf.DEPTH)
            stats = pbuf.scatter_stats(slot, g_pos, g_lse, nsum, n_local)      # (g_pos, g_lse, neg_sum) rows + barrier
            ws, npart = core.stats_bwd_packed(z, pbuf.z[slot], n_local, rank * n_local, n_local * w, temperature, g_pos,
                                              g_lse, nsum, stats, algo)
        else:
            if w > 1:
                packed = gather_global_order(torch.stack((g_pos, g_lse, nsum), dim=1), group)
                gp_c, gl_c, ns_c = (_contig(packed[:, k]) for k in range(3))
            else:
                gp_c, gl_c, ns_c = g_pos, g_lse, nsum
            ws, npart = core.stats_bwd(z, z_cols, n_local, rank * n_local, n_local * w, temperature, g_pos, g_lse, nsum,
                                       gp_c, gl_c, ns_c, algo)
        dp1, dp2 = core.normalize_bwd(ws, npart, 1.0, z, inv, n_local, n_local, p_dtype)
        return dp1, dp2, None, None, None, None


def cal_logits(p1: torch.Tensor, p2: torch.Tensor, temperature: float, precision: str = "auto", group=None):
    """Drop-in for the body of ``_cal_logits`` after the projectors (simclr.py:293-322).

    p1, p2: projector outputs ``[N, D]`` (first / second halves of the reference's ``torch.cat``).
    Returns ``(logits[2N, 2] fp32, labels[2N] int64 zeros)`` such that ``nn.CrossEntropyLoss()(logits, labels)``
    equals the reference's loss on its ``[2N, 2N-1]`` logits, with identical gradients into p1 / p2.
    With ``group`` (torch.distributed process group) the negatives are all-gathered across ranks; each rank
    returns the logits of its own 2N rows (mean over ranks of the per-rank CE == global-batch loss).
    """
    if p1.dim() != 2 or p1.shape != p2.shape:
        raise ValueError("cal_logits expects two [N, D] tensors of identical shape")
    require_cuda(p1, p2)
    if temperature <= 0:
        raise ValueError("temperature must be > 0")
    z_dtype, algo = pick_precision(p1, precision)
    logits = _InfoNCELogits.apply(p1, p2, float(temperature), z_dtype, algo, group)
    labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
    return logits, labels


# ------------------------------------------------------------------------------------------------------
# N2: fused projector tail (last Linear + affine-free BatchNorm + F.normalize) feeding K2 / K3 directly
# ------------------------------------------------------------------------------------------------------
class TailSpec:
    """One application of a projector's last two layers: rows ``h`` [R, K], the Linear's weight cast to h's dtype, and
    the BatchNorm module (statistics source + running-stat side effect)."""
    __slots__ = ("h", "w", "bn")

    def __init__(self, h, w, bn):
        self.h, self.w, self.bn = h, w, bn


def tail_supported(h: torch.Tensor, linear, bn) -> bool:
    """Can ``bn(linear(h))`` followed by F.normalize go through the fused tail?  16-bit CUDA rows (autocast), a plain
    bias-free nn.Linear, an affine-free BatchNorm1d / SyncBatchNorm, K % 64 == 0, D in {64, 128, 192, 256}."""
    import torch.nn as nn
    if not (h.is_cuda and h.dim() == 2 and h.dtype in (torch.float16, torch.bfloat16)):
        return False
    if type(linear) is not nn.Linear or linear.bias is not None:
        return False
    if not isinstance(bn, (nn.BatchNorm1d, nn.SyncBatchNorm)) or bn.affine:
        return False
    if bn.training and h.shape[0] < 2:
        return False
    return bool(lib().sm3_proj_tail_supported(h.shape[1], linear.weight.shape[0], dtype_code(h)))


def _bn_sync_group(bn):
    import torch.nn as nn
    if isinstance(bn, nn.SyncBatchNorm) and bn.training and dist.is_available() and dist.is_initialized():
        g = bn.process_group if bn.process_group is not None else dist.group.WORLD
        if dist.get_world_size(g) > 1:
            return g
    return None


class _TailInfoNCELogits(torch.autograd.Function):
    """[Linear -> BatchNorm(affine=False) -> normalise] for one (2N rows) or two (N rows each) row groups, then the K2
    statistics; backward = K3 -> normalise / BatchNorm backward kernels -> two cuBLAS GEMMs per group."""

    @staticmethod
    def forward(ctx, temperature, group, bns, *hw):
        k = len(bns)
        hs = [_contig(t) for t in hw[:k]]
        ws = [_contig(t) for t in hw[k:]]
        dev = require_cuda(*hs, *ws)
        d = ws[0].shape[0]
        rows = [h.shape[0] for h in hs]
        m = sum(rows)
        n_local = m // 2
        w_, rank = _group_info(group)
        z = torch.empty((m, d), dtype=torch.bfloat16, device=dev)
        inv = torch.empty(m, dtype=torch.float32, device=dev)
        saved, meta = [], []
        off = 0
        with torch.cuda.device(dev):
            for h, w, bn, r in zip(hs, ws, bns, rows):
                kdim = h.shape[1]
                y = torch.empty((r, d), dtype=torch.float32, device=dev)
                totals = torch.empty(2 * d, dtype=torch.float32, device=dev)
                wsb = torch.empty(int(lib().sm3_proj_tail_workspace_bytes(r, d)), dtype=torch.uint8, device=dev)
                training = bn.training or not bn.track_running_stats
                count = float(r)
                sync = _bn_sync_group(bn)
                check(lib().sm3_proj_tail_gemm(ptr(h), ptr(w), r, kdim, d, dtype_code(h), ptr(y), ptr(totals), ptr(wsb),
                                               wsb.numel(), stream_ptr()), "sm3_proj_tail_gemm")
                if training and sync is not None:              # SyncBatchNorm: global statistics (backbone_train.py:510)
                    dist.all_reduce(totals, group=sync)        # equal shards per rank (DistributedSampler, misc.py:400)
                    count = float(r * dist.get_world_size(sync))
                mom = 0.0
                rm = rv = None
                if bn.track_running_stats:
                    rm, rv = bn.running_mean, bn.running_var
                    if bn.training:
                        bn.num_batches_tracked += 1
                        mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                mean = torch.empty(d, dtype=torch.float32, device=dev)
                rstd = torch.empty(d, dtype=torch.float32, device=dev)
                upd = bool(bn.training and bn.track_running_stats)
                check(lib().sm3_proj_tail_bn_l2(ptr(y), r, d, ptr(totals), count, bn.eps, _EPS, int(training), mom,
                                                ptr(rm) if (upd or not training) else None,
                                                ptr(rv) if (upd or not training) else None, ptr(mean), ptr(rstd),
                                                z.data_ptr() + off * d * 2, inv.data_ptr() + off * 4, stream_ptr()),
                      "sm3_proj_tail_bn_l2")
                saved += [h, w, y, mean, rstd]
                meta.append((off, r, count, training, sync, h.dtype))
                off += r
        z_cols = gather_global_order(z, group) if w_ > 1 else z
        pos, lse, nsum = core.stats_fwd(z, z_cols, n_local, rank * n_local, n_local * w_, temperature, ALGO_TC)
        ctx.save_for_backward(z, inv, z_cols, nsum, *saved)
        ctx.meta = (temperature, group, w_, rank, n_local, d, meta)
        return torch.stack((pos, lse), dim=1)

    @staticmethod
    def backward(ctx, g):
        z, inv, z_cols, nsum, *saved = ctx.saved_tensors
        temperature, group, w_, rank, n_local, d, meta = ctx.meta
        dev = z.device
        g = g.float()
        g_pos, g_lse = _contig(g[:, 0]), _contig(g[:, 1])
        if w_ > 1:
            packed = gather_global_order(torch.stack((g_pos, g_lse, nsum), dim=1), group)
            gp_c, gl_c, ns_c = (_contig(packed[:, k]) for k in range(3))
        else:
            gp_c, gl_c, ns_c = g_pos, g_lse, nsum
        ws, npart = core.stats_bwd(z, z_cols, n_local, rank * n_local, n_local * w_, temperature, g_pos, g_lse, nsum,
                                   gp_c, gl_c, ns_c, ALGO_TC)
        m = 2 * n_local
        dhs, dws = [], []
        with torch.cuda.device(dev):
            for gi, (off, r, count, training, sync, h_dtype) in enumerate(meta):
                h, w, y, mean, rstd = saved[5 * gi: 5 * gi + 5]
                dyhat = torch.empty((r, d), dtype=torch.float32, device=dev)
                totals2 = torch.empty(2 * d, dtype=torch.float32, device=dev)
                wsb = torch.empty(int(lib().sm3_proj_tail_workspace_bytes(r, d)), dtype=torch.uint8, device=dev)
                check(lib().sm3_proj_tail_bwd1(ws.data_ptr() + off * d * 4, npart, m * d, z.data_ptr() + off * d * 2,
                                               inv.data_ptr() + off * 4, _EPS, ptr(y), ptr(mean), ptr(rstd), r, d, ptr(dyhat),
                                               ptr(totals2), ptr(wsb), wsb.numel(), stream_ptr()), "sm3_proj_tail_bwd1")
                if sync is not None:
                    dist.all_reduce(totals2, group=sync)
                dy = torch.empty((r, d), dtype=h_dtype, device=dev)
                check(lib().sm3_proj_tail_bwd2(ptr(dyhat), ptr(y), ptr(mean), ptr(rstd), ptr(totals2), count, int(training), r,
                                               d, ptr(dy), dtype_code(dy), stream_ptr()), "sm3_proj_tail_bwd2")
                dhs.append(dy @ w)                     # [R, D] x [D, K]   (cuBLAS)
                dws.append(dy.t() @ h)                 # [D, R] x [R, K]
        return (None, None, None) + tuple(dhs) + tuple(dws)


def tail_cal_logits(tails: Sequence[TailSpec], temperature: float, group=None):
    """``cal_logits`` with the projector tails fused in: ``tails`` is one TailSpec over all 2N rows (SimCLR.forward,
    simclr.py:61: BatchNorm over both views jointly) or two over N rows each (_cal_logits, :293: one BatchNorm pass per
    modality).  Returns ``(logits [2N, 2], zeros [2N])`` like ``cal_logits``; gradients flow to every ``h`` and ``w``."""
    if len(tails) not in (1, 2):
        raise ValueError("tail_cal_logits takes one (2N rows) or two (N rows each) row groups")
    rows = [t.h.shape[0] for t in tails]
    if (len(tails) == 1 and rows[0] % 2) or (len(tails) == 2 and rows[0] != rows[1]):
        raise ValueError("the row groups must split into two halves of N rows")
    logits = _TailInfoNCELogits.apply(float(temperature), group, tuple(t.bn for t in tails), *[t.h for t in tails],
                                      *[t.w for t in tails])
    labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
    return logits, labels


def _resolve_comm(comm: str, w: int, z_dtype, n_global: int, d: int, device, group):
    """-> PeerBuffers or None (NCCL).  'auto' uses the peer-memory exchange when symmetric memory can be set up."""
    if w == 1 or comm == "nccl" or z_dtype != torch.bfloat16:
        if comm == "peer" and w > 1:
            raise RuntimeError("comm='peer' needs the bf16 tensor-core path")
        return None
    from . import peer
    try:
        return peer.get_peer_buffers(group, n_global, d, device)
    except peer.PeerUnavailable:
        if comm == "peer":
            raise
        return None


def _fused_mode() -> int:
    """Exchange mode of sm3_infonce_step_peer for the fused path: 4 = symmetric forward across ranks (every rank computes
    half of its row block's column blocks and ships the column sums of the rest to their owners; SM3_PEER_SYM=0 disables),
    3 = rows pushed from inside K2 (SM3_PEER_PUSH=1), 2 = rows pushed by the normalise kernel."""
    if os.environ.get("SM3_PEER_PUSH", "0") == "1":
        return 3
    return 4 if os.environ.get("SM3_PEER_SYM", "1") != "0" else 2


class _FusedInfoNCE(torch.autograd.Function):
    """Scalar loss with the backward computed eagerly in forward (one pass: fwd + bwd kernels back to back)."""

    @staticmethod
    def forward(ctx, p1, p2, temperature, z_dtype, algo, group, weight, comm):
        w, rank = _group_info(group)
        n_local = p1.shape[0]
        n_global = n_local * w
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        ctx.in_dtypes = (p1.dtype, p2.dtype)
        p1, p2 = _grad_safe(p1), _grad_safe(p2)
        if w == 1 and _PROFILE is None:
            # single GPU: the whole step is enqueued by one C call (launch-bound shapes: ~8 launches, no Python
            # between them)
            p1c, p2c = _contig(p1), _contig(p2)
            dev, d = p1c.device, p1c.shape[1]
            io = dtype_code(p1c)
            with torch.cuda.device(dev):
                scratch = _scratch("step", dev, (n_local, d, io, algo),
                                   lambda: lib().sm3_infonce_step_scratch_bytes(n_local, d, io, algo))
                loss = torch.empty((), dtype=torch.float32, device=dev)
                dp1 = torch.empty_like(p1c) if need_grad else None
                dp2 = torch.empty_like(p2c) if need_grad else None
                check(lib().sm3_infonce_step(ptr(p1c), ptr(p2c), n_local, d, io, temperature, weight, ptr(loss),
                                             ptr(dp1), ptr(dp2), ptr(scratch), scratch.numel(), algo, stream_ptr()),
                      "sm3_infonce_step")
            if need_grad:
                ctx.save_for_backward(dp1, dp2)
            ctx.comm_used = "none"
            return loss
        pbuf = _resolve_comm(comm, w, z_dtype, n_global, p1.shape[1], p1.device, group)
        slot = pbuf.next_slot() if pbuf is not None else 0
        _mark("start")
        z, inv = core.normalize_pair(p1, p2, z_dtype)
        _mark("normalize")
        # Overlapping the exchange with the local column block (SM3_PEER_OVERLAP=1) is implemented and verified, but on
        # B200 it does not pay yet: splitting K2/K3 into a local and a remote launch costs about what the hidden
        # exchange saves (8 ranks: 0.754 ms overlapped vs 0.703 ms back to back), so it is opt-in.
        ov_env = os.environ.get("SM3_PEER_OVERLAP")
        overlap = (pbuf is not None and n_local % 128 == 0 and algo == ALGO_TC and
                   ov_env == "1")
        # Fused exchange (sm3_infonce_step_peer mode 2, default; SM3_PEER_FUSED=0 disables): the producers scatter +
        # signal themselves and K2 / K3 wait inside the kernel, 5 launches per step instead of 13.  Measured on B200,
        # cfg4: 2 ranks 2.563 -> 2.493 ms, 8 ranks 0.778 -> 0.701 ms per step.
        fused = (pbuf is not None and n_local % 128 == 0 and algo == ALGO_TC and not overlap and
                 os.environ.get("SM3_PEER_FUSED", "1") != "0")
        if pbuf is not None and algo == ALGO_TC and _PROFILE is None and pbuf.multicast is False:
            # ---- the whole multi-rank step enqueued by one C call (exchange overlapped or back to back) ----
            p1c, p2c = _contig(p1), _contig(p2)
            dev, d = p1c.device, p1c.shape[1]
            main = torch.cuda.current_stream()
            with torch.cuda.device(dev):
                scratch = _scratch("peer", dev, (n_local, n_global, d),
                                   lambda: lib().sm3_infonce_step_peer_scratch_bytes(n_local, n_global, d))
                loss = torch.empty((), dtype=torch.float32, device=dev)
                dp1 = torch.empty_like(p1c) if need_grad else None
                dp2 = torch.empty_like(p2c) if need_grad else None
                scratch.record_stream(pbuf.side)
                # mode 2 = fused exchange, rows pushed by the normalise kernel (default); SM3_PEER_PUSH=1 -> mode 3: rows
                # pushed from inside K2 with owner-ordered tiles.  Measured on 8 x B200 (cfg4, profiles/r02_scale8_*.json):
                # mode 3 shortens the kernels (stage sum 0.69-0.72 vs 0.81 ms) but not the free-running step (0.70-0.74
                # vs 0.74 ms) and is 7 % slower at 2 ranks, so it stays opt-in.
                mode = _fused_mode() if fused else int(overlap)
                check(lib().sm3_infonce_step_peer(ptr(p1c), ptr(p2c), n_local, rank, w, d, dtype_code(p1c), temperature,
                                                  weight, ptr(loss), ptr(dp1), ptr(dp2), ptr(pbuf.z[slot]),
                                                  pbuf.zp[slot], ptr(pbuf.st[slot]), pbuf.stp[slot], ptr(pbuf.flags),
                                                  pbuf.fp, pbuf.step & 0x7FFFFFFF, mode, ptr(scratch),
                                                  scratch.numel(), main.cuda_stream, pbuf.side.cuda_stream),
                      "sm3_infonce_step_peer")
            if need_grad:
                ctx.save_for_backward(dp1, dp2)
            ctx.comm_used = ({3: "peer-fused-push", 4: "peer-fused-sym"}.get(mode, "peer-fused")) if fused else ("peer-overlap" if overlap else "peer")
            return loss
        if overlap:
            # ---- exchange on a side stream, local column block on the main stream, then the remote blocks ----
            main = torch.cuda.current_stream()
            epoch = pbuf.step                                   # next_slot() already advanced it: unique, growing
            off = rank * n_local
            pbuf.side.wait_stream(main)
            z.record_stream(pbuf.side)
            with torch.cuda.stream(pbuf.side):
                pbuf.store_z(slot, z, n_local)                  # NVLink stores into every rank's z_cols[slot]
                pbuf.signal(0, epoch)
            pos, lse_l, ns_l = core.stats_fwd(z, z, n_local, 0, n_local, temperature, algo)   # local block
            _mark("stats_fwd_local")
            main.wait_stream(pbuf.side)
            pbuf.wait(0, epoch)
            _mark("gather_z")
            z_cols = pbuf.z[slot]
            lse, nsum = core.stats_fwd_remote(z, z_cols, n_local, off, n_global, temperature, ns_l, pos)
            _mark("stats_fwd")
            loss, g_pos, g_lse = core.loss(pos, lse, weight / (2 * n_local), want_grads=need_grad)
            _mark("loss")
            if need_grad:
                pbuf.side.wait_stream(main)
                for t in (g_pos, g_lse, nsum):
                    t.record_stream(pbuf.side)
                with torch.cuda.stream(pbuf.side):
                    pbuf.store_stats(slot, g_pos, g_lse, nsum, n_local)
                    pbuf.signal(1, epoch)

                def _between():
                    _mark("stats_bwd_local")
                    main.wait_stream(pbuf.side)
                    pbuf.wait(1, epoch)
                    _mark("gather_stats")

                ws, npart = core.stats_bwd_split(z, z_cols, n_local, off, n_global, temperature, g_pos, g_lse, nsum,
                                                 pbuf.st[slot], _between)
                _mark("stats_bwd")
                dp1, dp2 = core.normalize_bwd(ws, npart, 1.0, z, inv, n_local, n_local, p1.dtype)
                _mark("normalize_bwd")
                ctx.save_for_backward(dp1, dp2)
            ctx.comm_used = "peer-overlap"
            return loss
        if w == 1:
            z_cols = z
        elif pbuf is not None:
            z_cols = pbuf.scatter_z(slot, z, n_local)          # NVLink peer stores + cross-rank barrier
        else:
            z_cols = gather_global_order(z, group)            # NCCL all-gather
        _mark("gather_z")
        pos, lse, nsum = core.stats_fwd(z, z_cols, n_local, rank * n_local, n_global, temperature, algo)
        _mark("stats_fwd")
        loss, g_pos, g_lse = core.loss(pos, lse, weight / (2 * n_local), want_grads=need_grad)
        _mark("loss")
        if need_grad:
            if pbuf is not None:
                stats = pbuf.scatter_stats(slot, g_pos, g_lse, nsum, n_local)
                _mark("gather_stats")
                ws, npart = core.stats_bwd_packed(z, z_cols, n_local, rank * n_local, n_global, temperature, g_pos,
                                                  g_lse, nsum, stats, algo)
            else:
                if w > 1:
                    packed = gather_global_order(torch.stack((g_pos, g_lse, nsum), dim=1), group)
                    gp_c, gl_c, ns_c = (_contig(packed[:, k]) for k in range(3))
                else:
                    gp_c, gl_c, ns_c = g_pos, g_lse, nsum
                _mark("gather_stats")
                ws, npart = core.stats_bwd(z, z_cols, n_local, rank * n_local, n_global, temperature, g_pos, g_lse,
                                           nsum, gp_c, gl_c, ns_c, algo)
            _mark("stats_bwd")
            dp1, dp2 = core.normalize_bwd(ws, npart, 1.0, z, inv, n_local, n_local, p1.dtype)
            _mark("normalize_bwd")
            ctx.save_for_backward(dp1, dp2)
        ctx.comm_used = "peer" if pbuf is not None else ("nccl" if w > 1 else "none")
        return loss

    @staticmethod
    def backward(ctx, g):
        o1, o2 = _scale_grads(ctx.saved_tensors, g, ctx.in_dtypes)
        return o1, o2, None, None, None, None, None, None


def fused_infonce(p1: torch.Tensor, p2: torch.Tensor, temperature: float, precision: str = "auto", group=None,
                  weight: float = 1.0, comm: str = "auto") -> torch.Tensor:
    """``weight * CrossEntropy(_cal_logits(p1, p2, T))`` as one fused op (fp32 scalar).

    Per-rank normalisation follows the reference / DDP convention: mean over this rank's 2N rows.
    ``comm`` (multi-rank only): 'peer' = NVLink peer-memory exchange (symmetric memory), 'nccl' = NCCL all-gather,
    'auto' = peer when it can be set up, else NCCL."""
    if p1.dim() != 2 or p1.shape != p2.shape:
        raise ValueError("fused_infonce expects two [N, D] tensors of identical shape")
    require_cuda(p1, p2)
    if temperature <= 0:
        raise ValueError("temperature must be > 0")
    z_dtype, algo = pick_precision(p1, precision)
    if comm not in ("auto", "peer", "nccl"):
        raise ValueError("comm must be 'auto', 'peer' or 'nccl'")
    return _FusedInfoNCE.apply(p1, p2, float(temperature), z_dtype, algo, group, float(weight), comm)


class _FusedInfoNCEMulti(torch.autograd.Function):
    """Several same-shape terms through sm3_infonce_step_multi: one C call, one device scalar."""

    @staticmethod
    def forward(ctx, temperature, algo, weights, *ps):
        import ctypes as C
        t = len(ps) // 2
        ctx.in_dtypes = tuple(p.dtype for p in ps)
        p1s = [_contig(_grad_safe(p)) for p in ps[:t]]
        p2s = [_contig(_grad_safe(p)) for p in ps[t:]]
        dev = require_cuda(*p1s, *p2s)
        n, d = p1s[0].shape
        io = dtype_code(p1s[0])
        need_grad = any(ctx.needs_input_grad[3:])
        with torch.cuda.device(dev):
            scratch = _scratch("multi", dev, (n, d, io, algo),
                               lambda: lib().sm3_infonce_step_multi_scratch_bytes(n, d, io, algo))
            loss = torch.empty((), dtype=torch.float32, device=dev)
            d1 = [torch.empty_like(p) for p in p1s] if need_grad else None
            d2 = [torch.empty_like(p) for p in p2s] if need_grad else None
            arr = lambda ts: (C.c_void_p * t)(*[x.data_ptr() for x in ts])       # noqa: E731
            check(lib().sm3_infonce_step_multi(t, arr(p1s), arr(p2s), n, d, io, temperature, (C.c_float * t)(*weights),
                                               ptr(loss), arr(d1) if need_grad else None, arr(d2) if need_grad else None,
                                               ptr(scratch), scratch.numel(), algo, stream_ptr()),
                  "sm3_infonce_step_multi")
        if need_grad:
            ctx.save_for_backward(*d1, *d2)
        return loss

    @staticmethod
    def backward(ctx, g):
        return (None, None, None) + _scale_grads(ctx.saved_tensors, g, ctx.in_dtypes)


def fused_infonce_multi(pairs, temperature: float, weights: Optional[Sequence[float]] = None,
                        precision: str = "auto") -> torch.Tensor:
    """``sum_t weights[t] * CrossEntropy(_cal_logits(p1_t, p2_t, T))`` for several same-shape terms in one call -- the
    derm / clinic / cross / cross terms of SimCLRSkinV3 style 0 (``loss = derm + clinic + 0.5*cross1 + 0.5*cross2``,
    tools/backbone_train.py:101-121).  ``pairs``: sequence of ``(p1, p2)`` projector outputs ``[N, D]``.  Single
    process (each rank its own negatives); the sharded form is ``fused_infonce(group=...)`` per term."""
    pairs = list(pairs)
    if not pairs:
        raise ValueError("fused_infonce_multi needs at least one (p1, p2) pair")
    shape, dtype = pairs[0][0].shape, pairs[0][0].dtype
    for p1, p2 in pairs:
        if p1.dim() != 2 or p1.shape != shape or p2.shape != shape or p1.dtype != dtype or p2.dtype != dtype:
            raise ValueError("fused_infonce_multi expects [N, D] tensors of one shape and dtype")
    if temperature <= 0:
        raise ValueError("temperature must be > 0")
    require_cuda(*[p for pr in pairs for p in pr])
    _, algo = pick_precision(pairs[0][0], precision)
    w = [1.0] * len(pairs) if weights is None else [float(x) for x in weights]
    if len(w) != len(pairs):
        raise ValueError("one weight per term")
    return _FusedInfoNCEMulti.apply(float(temperature), algo, tuple(w), *[p[0] for p in pairs], *[p[1] for p in pairs])


# ------------------------------------------------------------------------------------------------------
# heads
# ------------------------------------------------------------------------------------------------------
class _MultiHeadCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, class_counts, weights, inv_t, ignore_index, use_ignore):
        import ctypes as C
        dev = require_cuda(logits, labels)
        ctx.in_dtype = logits.dtype
        if logits.dtype == torch.float16:
            # fp16 cannot hold d loss / d logit ~ 1 / (H B) before the upstream (GradScaler) factor is applied: compute
            # in fp32 like the reference does under autocast (CE is an fp32 op there) and round once, after scaling
            logits = logits.float()
        logits = _contig(logits)
        labels = _contig(labels.to(torch.int64))
        b, c_total = logits.shape
        h = len(class_counts)
        if sum(class_counts) != c_total or labels.shape != (b, h):
            raise ValueError(f"logits [B,{c_total}] / labels {tuple(labels.shape)} do not match heads {class_counts}")
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dlogits = torch.empty_like(logits) if need_grad else None
        nbytes = lib().sm3_multihead_ce_workspace_bytes(b, h)
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        cc = (C.c_int * h)(*class_counts)
        wt = (C.c_float * h)(*weights) if weights is not None else None
        with torch.cuda.device(dev):
            check(lib().sm3_multihead_ce(ptr(logits), dtype_code(logits), ptr(labels), b, h, cc, wt, inv_t,
                                         int(use_ignore), ignore_index, ptr(loss), ptr(dlogits), 1.0, ptr(ws),
                                         ws.numel(), stream_ptr()), "sm3_multihead_ce")
        if need_grad:
            ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return (dlogits * g.to(dlogits.dtype)).to(ctx.in_dtype), None, None, None, None, None, None


def multihead_ce(outputs, labels: torch.Tensor, weights: Optional[Sequence[float]] = None, temperature: float = 1.0,
                 ignore_index: Optional[int] = None, class_counts: Optional[Sequence[int]] = None) -> torch.Tensor:
    """``sum_h w_h * CE(outputs[h] / T, labels[:, h]) / H`` in one launch (forward + backward).

    outputs: list of H tensors ``[B, n_h]`` (the reference heads' return value) or one ``[B, sum n_h]`` tensor.
    ignore_index=None promises there are no ignored labels (mlc_eval form); an int enables the
    DeepCluster form of tools/mlc_train.py:381 (``ignore_index=-100``)."""
    if isinstance(outputs, (list, tuple)):
        class_counts = [int(o.shape[1]) for o in outputs]
        logits = torch.cat(list(outputs), dim=1)
    else:
        logits = outputs
        class_counts = list(class_counts if class_counts is not None else NUM_CLASSES)
    w = None if weights is None else [float(x) for x in weights]
    return _MultiHeadCE.apply(logits, labels, tuple(class_counts), w, 1.0 / float(temperature),
                              -100 if ignore_index is None else int(ignore_index), ignore_index is not None)


class _BCEWithLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t, pos_weight):
        dev = require_cuda(x, t, pos_weight)
        ctx.in_dtype = x.dtype
        if x.dtype == torch.float16:
            x = x.float()              # see _MultiHeadCE: the 1/(B C) gradient is rounded to fp16 only after upstream scaling
        x = _contig(x)
        t = _contig(t)
        if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            t = t.to(torch.float32)
        if x.shape != t.shape or x.dim() != 2:
            raise ValueError("bce_with_logits expects logits and targets of identical shape [B, C]")
        b, c = x.shape
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x) if need_grad else None
        pw = None if pos_weight is None else _contig(pos_weight.float())
        ws = torch.empty(int(lib().sm3_bce_workspace_bytes(b, c)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().sm3_bce_logits(ptr(x), dtype_code(x), ptr(t), dtype_code(t), ptr(pw), b, c, ptr(loss), ptr(dx),
                                       1.0, ptr(ws), ws.numel(), stream_ptr()), "sm3_bce_logits")
        if need_grad:
            ctx.save_for_backward(dx)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return (dx * g.to(dx.dtype)).to(ctx.in_dtype), None, None


def bce_with_logits(x: torch.Tensor, target: torch.Tensor, pos_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean-reduced ``binary_cross_entropy_with_logits`` over multi-hot targets, forward + backward in one launch."""
    return _BCEWithLogits.apply(x, target, pos_weight)


# ------------------------------------------------------------------------------------------------------
# retrieval (N1)
# ------------------------------------------------------------------------------------------------------
def sim_topk(query: torch.Tensor, bank: torch.Tensor, k: int, exclude_self_offset: int = -1):
    """``(query @ bank.T).topk(k, dim=-1)`` without materialising the similarity matrix
    (reference src/models/evaluator.py:61-63).  Returns (values fp32 [B,k] descending, indices int64 [B,k]);
    ties resolve to the lower bank index.  ``exclude_self_offset >= 0`` masks bank row (offset + q) for query q
    (in-batch retrieval of the non-self neighbour, tools/backbone_train.py:103-105)."""
    dev = require_cuda(query, bank)
    if query.dim() != 2 or bank.dim() != 2 or query.shape[1] != bank.shape[1] or query.dtype != bank.dtype:
        raise ValueError("sim_topk expects query [B, D] and bank [N, D] of the same dtype")
    query, bank = _contig(query), _contig(bank)
    b, d = query.shape
    vals = torch.empty((b, k), dtype=torch.float32, device=dev)
    idx = torch.empty((b, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        # SM3_TOPK_TILED: 0 single pass | 1 tiled + threshold filter | 2 materialised + radix select | unset: 2 unless the
        # bank is tiny (one CTA per query only pays off with >= ~1000 bank rows to scan).  Measured on B200, us: 512 x 16384 x
        # 128, k = 200: 1582 / 628 / 176 (library 250); 8192 x 8192 x 128, k = 5: 6860 / 1341 / 934 (library 1280).
        form = os.environ.get("SM3_TOPK_TILED") or ("2" if bank.shape[0] >= 1024 else "1")
        if form == "2":                             # materialised similarities + radix select per query
            nbytes = int(lib().sm3_sim_topk_mat_workspace_bytes(b, bank.shape[0], int(k)))
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            check(lib().sm3_sim_topk_mat(ptr(query), ptr(bank), b, bank.shape[0], d, dtype_code(query), int(k),
                                         int(exclude_self_offset), ptr(vals), ptr(idx), ptr(ws), ws.numel(),
                                         stream_ptr()), "sm3_sim_topk_mat")
        elif form != "0":
            nbytes = int(lib().sm3_sim_topk_workspace_bytes(b, bank.shape[0], int(k)))
            ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
            check(lib().sm3_sim_topk_ws(ptr(query), ptr(bank), b, bank.shape[0], d, dtype_code(query), int(k),
                                        int(exclude_self_offset), ptr(vals), ptr(idx), ptr(ws), ws.numel(),
                                        stream_ptr()), "sm3_sim_topk_ws")
        else:                                       # single-pass form (no workspace), kept for comparison
            check(lib().sm3_sim_topk(ptr(query), ptr(bank), b, bank.shape[0], d, dtype_code(query), int(k),
                                     int(exclude_self_offset), ptr(vals), ptr(idx), stream_ptr()), "sm3_sim_topk")
    return vals, idx


def knn_predict(query: torch.Tensor, bank: torch.Tensor, bank_labels: torch.Tensor, num_classes: int, k: int = 200,
                temperature: float = 0.07) -> torch.Tensor:
    """Weighted k-NN vote of KNNOnlineEvaluator.predict (reference src/models/evaluator.py:43-83): class ids per
    query sorted by descending score.  The top-k search is the fused kernel; the k-way vote is a [B, k] scatter."""
    w, idx = sim_topk(query, bank, k)
    labels = bank_labels.to(idx.device)[idx]                                   # [B, k]
    w = (w / temperature).exp()
    scores = torch.zeros((query.shape[0], num_classes), dtype=torch.float32, device=idx.device)
    scores.scatter_add_(1, labels, w)
    return scores.argsort(dim=-1, descending=True)


# ------------------------------------------------------------------------------------------------------
# N3: prototype heads of the multi-label block (tail of Model.forward, reference tools/mlc_train.py:70-89)
# ------------------------------------------------------------------------------------------------------
_PROTO_INDEX = {}


def _proto_index(slots, hf: int, dev: torch.device):
    """Cached index tensors for the batched dW GEMM of the prototype heads: `gather` picks, per feature slot, the logit
    columns that read it (padded with the index of an appended zero column), `scatter` maps the [hf * width] result rows
    back to class order."""
    key = (tuple(slots), hf, dev.index)
    hit = _PROTO_INDEX.get(key)
    if hit is None:
        per = [[c for c, s_ in enumerate(slots) if s_ == slot] for slot in range(hf)]
        width = max(len(p) for p in per)
        c_total = len(slots)
        gather = [c for p in per for c in (p + [c_total] * (width - len(p)))]
        where = {c: i for i, c in enumerate(gather) if c < c_total}
        scatter = [where[c] for c in range(c_total)]
        hit = (torch.tensor(gather, device=dev), torch.tensor(scatter, device=dev), width)
        _PROTO_INDEX[key] = hit
    return hit


class _ProtoHeads(torch.autograd.Function):
    """(normalised feats, [B, C] logits) from sm3_proto_heads_fwd; backward = sm3_proto_heads_bwd + one GEMM for dW."""

    @staticmethod
    def forward(ctx, feats, w_cat, slots, l2_norm):
        import ctypes as C
        dev = require_cuda(feats, w_cat)
        hf, b, d = feats.shape
        c = w_cat.shape[0]
        feats = _contig(feats)
        w32 = _contig(w_cat.float())
        logits = torch.empty((b, c), dtype=torch.float32, device=dev)
        z = torch.empty_like(feats) if l2_norm else feats
        inv = torch.empty(hf * b, dtype=torch.float32, device=dev) if l2_norm else None
        slot_arr = (C.c_int * c)(*slots)
        with torch.cuda.device(dev):
            check(lib().sm3_proto_heads_fwd(ptr(feats), dtype_code(feats), hf, b, d, ptr(w32), c, slot_arr, int(l2_norm),
                                            1e-12, ptr(z) if l2_norm else None, ptr(inv), ptr(logits), stream_ptr()),
                  "sm3_proto_heads_fwd")
        ctx.save_for_backward(z, w32, inv if inv is not None else torch.empty(0, device=dev))
        ctx.slots, ctx.l2_norm, ctx.w_dtype = tuple(slots), bool(l2_norm), w_cat.dtype
        ctx.set_materialize_grads(False)          # the returned features usually carry no gradient (memory bank only)
        return z, logits

    @staticmethod
    def backward(ctx, dz_extra, dlogits):
        import ctypes as C
        z, w32, inv = ctx.saved_tensors
        hf, b, d = z.shape
        c = w32.shape[0]
        dlogits = torch.zeros((b, c), dtype=torch.float32, device=z.device) if dlogits is None else _contig(dlogits.float())
        d_feats = dw = None
        slot_arr = (C.c_int * c)(*ctx.slots)
        if ctx.needs_input_grad[0]:
            d_feats = torch.empty_like(z)
            extra = _contig(dz_extra.to(z.dtype)) if dz_extra is not None else None
            with torch.cuda.device(z.device):
                check(lib().sm3_proto_heads_bwd(ptr(z), dtype_code(z), hf, b, d, ptr(w32), c, slot_arr, int(ctx.l2_norm),
                                                ptr(inv) if ctx.l2_norm else None, ptr(dlogits), ptr(extra), ptr(d_feats),
                                                stream_ptr()), "sm3_proto_heads_bwd")
        if ctx.needs_input_grad[1]:
            # dW[c] = sum_b dlogit[b, c] * z[slot(c), b, :]  (no host synchronisation: index tensors are cached)
            if hf == 1:
                dw = dlogits.t() @ z[0].float()
            else:
                gather, scatter, width = _proto_index(ctx.slots, hf, z.device)
                dl = torch.cat([dlogits, dlogits.new_zeros(b, 1)], dim=1)[:, gather].view(b, hf, width)
                per_slot = torch.bmm(dl.permute(1, 2, 0), z.float())                  # [hf, width, d]
                dw = per_slot.reshape(hf * width, d)[scatter]
            dw = dw.to(ctx.w_dtype)
        return d_feats, dw, None, None


def proto_heads_supported(feats: torch.Tensor, n_classes_total: int) -> bool:
    return feats.is_cuda and feats.dim() == 3 and feats.dtype in (torch.float32, torch.float16, torch.bfloat16) and \
        bool(lib().sm3_proto_heads_supported(int(feats.shape[2]), int(n_classes_total), dtype_code(feats)))


def proto_heads(sa_feats: torch.Tensor, prototype_weights: Sequence[torch.Tensor], l2_norm: bool = False):
    """The tail of the multi-label ``Model.forward`` (reference tools/mlc_train.py:81-87) in one launch: optional
    ``F.normalize(sa_feats[i], dim=-1)`` of every feature slot and ``preds[i] = prototypes[i](sa_feats[i % len(sa_feats)])``.
    ``sa_feats`` [Hf, B, D]; ``prototype_weights`` = the H bias-free Linear weights ``[n_i, D]``.  Returns
    ``(sa_feats_out [Hf, B, D], logits [B, sum n_i] fp32)``: ``logits.split(n_i, 1)`` are the reference's ``preds`` and the
    tensor is what ``multihead_ce(logits, targets, class_counts=n_i, ...)`` consumes without a ``torch.cat``."""
    hf = sa_feats.shape[0]
    counts = [int(w.shape[0]) for w in prototype_weights]
    slots = [h % hf for h, n in enumerate(counts) for _ in range(n)]
    w_cat = torch.cat(list(prototype_weights), dim=0)
    return _ProtoHeads.apply(sa_feats, w_cat, slots, bool(l2_norm))


def mlc_model_forward(self, derm_imgs, clinic_imgs):
    """Drop-in body for ``Model.forward`` of tools/mlc_train.py:70-89 (same attributes, same return structure
    ``(sa_feats, preds)``): everything up to the self-attention layer is the reference's own sequence, the normalise +
    eight prototype Linears run as the fused kernel.  Falls back to the stock tail for shapes the kernel does not take."""
    feats = self.extractor.extract(derm_imgs, clinic_imgs)
    feats = torch.cat(feats, dim=1)
    proj_feats = self.projectors(feats)
    if not isinstance(proj_feats, list):
        proj_feats = [proj_feats]
    proj_feats = torch.stack(proj_feats, dim=0)
    sa_feats = self.mlc_sa(proj_feats)
    weights = [p.weight for p in self.prototypes]
    total = sum(int(w.shape[0]) for w in weights)
    if not proto_heads_supported(sa_feats, total):
        if self.l2_norm:
            sa_feats = torch.nn.functional.normalize(sa_feats, dim=-1, p=2)
        return sa_feats, [p(sa_feats[i % len(sa_feats)]) for i, p in enumerate(self.prototypes)]
    sa_out, logits = proto_heads(sa_feats, weights, self.l2_norm)
    preds = list(torch.split(logits.to(sa_feats.dtype), [int(w.shape[0]) for w in weights], dim=1))
    return sa_out, preds


# ------------------------------------------------------------------------------------------------------
# DeepCluster memory-bank clustering (N4)
# ------------------------------------------------------------------------------------------------------
class _kmeans_core:
    """The two fused k-means kernels (csrc/kmeans.cu); tests/dist_worker.py swaps in a CPU stand-in for the gloo run."""

    @staticmethod
    def supported(d: int, k: int, emb: torch.Tensor) -> bool:
        return emb.dtype == torch.float32 and bool(lib().sm3_kmeans_supported(d, k, _lib.F32))

    @staticmethod
    def assign(emb: torch.Tensor, cent: torch.Tensor, want_sums: bool):
        """-> (assign int64 [n], packed [K*D + K] fp32 = cluster sums then counts, or None)"""
        dev = require_cuda(emb, cent)
        n, d = emb.shape
        k = cent.shape[0]
        assign = torch.empty(n, dtype=torch.int64, device=dev)
        packed = torch.empty(k * d + k, dtype=torch.float32, device=dev) if want_sums else None
        with torch.cuda.device(dev):
            ws = _scratch("kmeans", dev, (n, d, k), lambda: lib().sm3_kmeans_workspace_bytes(n, d, k)) if want_sums else None
            check(lib().sm3_kmeans_assign(ptr(emb), n, d, ptr(cent), k, ptr(assign), ptr(packed),
                                          (packed.data_ptr() + 4 * k * d) if want_sums else None, ptr(ws),
                                          ws.numel() if want_sums else 0, stream_ptr()), "sm3_kmeans_assign")
        return assign, packed

    @staticmethod
    def update(packed: torch.Tensor, cent: torch.Tensor) -> torch.Tensor:
        dev = require_cuda(packed, cent)
        k, d = cent.shape
        new = torch.empty_like(cent)
        with torch.cuda.device(dev):
            check(lib().sm3_kmeans_update(ptr(packed), packed.data_ptr() + 4 * k * d, ptr(cent), ptr(new), d, k, _EPS,
                                          stream_ptr()), "sm3_kmeans_update")
        return new


kmeans_core = _kmeans_core


@torch.no_grad()
def spherical_kmeans(emb: torch.Tensor, init_idx: Optional[torch.Tensor], n_iters: int = 10, group=None,
                     init_centroids: Optional[torch.Tensor] = None):
    """The rank-0 loop of ``cluster_memory`` (reference tools/mlc_train.py:144-176) without leaving the GPU: centroids
    start at ``emb[init_idx]`` (or ``init_centroids``); per iteration ONE pass over the bank does the E step (argmax of
    ``emb @ centroids.T``, never materialised, ties to the lower centroid) and accumulates the M step's per-cluster sums
    (``sm3_kmeans_assign``, deterministic), then ``sm3_kmeans_update`` takes the means of the non-empty clusters and
    L2-normalises all centroids -- no ``.cpu().numpy()`` / scipy.sparse / Python loop over clusters as at :161-172.
    With ``group`` the bank stays SHARDED: ``emb`` is this rank's shard, the [K*D + K] sums / counts are all-reduced
    between the two kernels and every rank ends with identical centroids.  Shapes the fused kernels do not take
    (D % 128 != 0, D > 512, K > 8) use ``sm3_sim_topk`` (k = 1) + a one-hot GEMM instead.
    Returns (assignments int64 [n], centroids fp32 [K, D])."""
    require_cuda(emb, init_idx, init_centroids)
    if emb.dim() != 2:
        raise ValueError("spherical_kmeans expects emb [n, D]")
    emb = _contig(emb.float())
    cent = _contig(init_centroids.float().clone()) if init_centroids is not None else _contig(emb[init_idx.long()].clone())
    k, d = cent.shape
    w, _ = _group_info(group)
    if kmeans_core.supported(d, k, emb):
        assign = None
        for it in range(n_iters + 1):
            last = it == n_iters
            assign, packed = kmeans_core.assign(emb, cent, not last)
            if last:
                break
            if w > 1:
                dist.all_reduce(packed, group=group)
            cent = kmeans_core.update(packed, cent)
        return assign, cent
    assign = None
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False          # the cluster sums must be fp32 sums whatever the script set
    try:
        for it in range(n_iters + 1):
            _, idx = sim_topk(emb, cent, 1)
            assign = idx[:, 0]
            if it == n_iters:
                break
            onehot = torch.nn.functional.one_hot(assign, k).to(emb.dtype)        # [n, K]
            counts = onehot.sum(0)                                               # exact in fp32 below 2^24 samples
            sums = onehot.t() @ emb                                              # [K, D]
            if w > 1:
                packed = torch.cat([sums.reshape(-1), counts])
                dist.all_reduce(packed, group=group)
                sums, counts = packed[: k * d].view(k, d), packed[k * d:]
            keep = (counts > 0).unsqueeze(1)
            cent = torch.where(keep, sums / counts.clamp(min=1.0).unsqueeze(1), cent)
            cent = l2_normalize(_contig(cent))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return assign, cent


@torch.no_grad()
def cluster_memory(args, prototypes, K: int, local_memory_index: torch.Tensor, local_memory_embeddings: torch.Tensor,
                   nmb_kmeans_iters: int = 10) -> torch.Tensor:
    """Drop-in for ``cluster_memory`` of tools/mlc_train.py:116-189 (same signature, same return value, same side
    effect ``prototypes.weight.copy_(centroids)``).  The bank stays sharded: every rank clusters its own
    ``local_memory_embeddings`` against shared centroids and only [K*D + K] floats per iteration cross ranks
    (all-reduce), instead of gathering the whole bank to rank 0, looping in Python there and broadcasting the result
    (:137-143, :185-186).  The K initial centroids are the rows rank 0's ``torch.randperm`` picks out of the rank-major
    concatenation of the shards, exactly as at :144-145."""
    dev = require_cuda(local_memory_embeddings)
    index = _contig(local_memory_index.to(dev).long())
    emb = _contig(local_memory_embeddings.float())
    world = int(getattr(args, "world_size", 1))
    rank = int(getattr(args, "rank", 0))
    n_local, d = emb.shape
    n = n_local * world
    if n < K:
        raise ValueError("please reduce the number of centroids")          # reference :146
    init = torch.randperm(n)[:K].to(dev) if rank == 0 else torch.empty(K, dtype=torch.long, device=dev)
    group = None
    if world > 1:
        dist.broadcast(init, 0)
        group = dist.group.WORLD
        mine = (init >= rank * n_local) & (init < (rank + 1) * n_local)
        cent0 = torch.zeros((K, d), dtype=torch.float32, device=dev)
        cent0[mine] = emb[(init[mine] - rank * n_local)]
        dist.all_reduce(cent0)                                              # each row has exactly one owner
    else:
        cent0 = emb[init]
    assign, centroids = spherical_kmeans(emb, None, nmb_kmeans_iters, group=group, init_centroids=cent0)
    if world > 1:
        all_assign = torch.empty(n, dtype=assign.dtype, device=dev)
        all_idx = torch.empty(n, dtype=index.dtype, device=dev)
        dist.all_gather_into_tensor(all_assign, _contig(assign))
        dist.all_gather_into_tensor(all_idx, index)
    else:
        all_assign, all_idx = assign, index
    assignments = torch.full((n,), -100, dtype=torch.long, device=dev)
    assignments[all_idx] = all_assign
    prototypes.weight.copy_(centroids.to(prototypes.weight.dtype))
    return assignments


# ------------------------------------------------------------------------------------------------------
# host-buffer entry (the call bench.py times end to end)
# ------------------------------------------------------------------------------------------------------
class HostInfoNCE:
    """Pinned-host-buffer front end of ``sm3_infonce_host``: H2D copy, fused fwd+bwd, D2H of loss and grads."""

    def __init__(self, n_pairs: int, d: int, dtype: torch.dtype = torch.bfloat16, algo: int = ALGO_AUTO,
                 device: Optional[torch.device] = None):
        self.n, self.d, self.dtype, self.algo = n_pairs, d, dtype, algo
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        code = _lib._DTYPES[dtype]
        with torch.cuda.device(self.device):
            nbytes = lib().sm3_infonce_host_scratch_bytes(n_pairs, d, code, algo)
        if nbytes == 0:
            raise ValueError("bad shape for HostInfoNCE")
        self.scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.loss = torch.empty(1, dtype=torch.float32).pin_memory()
        self.dp1 = torch.empty((n_pairs, d), dtype=dtype).pin_memory()
        self.dp2 = torch.empty((n_pairs, d), dtype=dtype).pin_memory()
        self.h2d_bytes = 2 * n_pairs * d * self.dp1.element_size()
        self.d2h_bytes = 4 + self.h2d_bytes

    def __call__(self, p1_host: torch.Tensor, p2_host: torch.Tensor, temperature: float):
        assert not p1_host.is_cuda and p1_host.dtype == self.dtype and tuple(p1_host.shape) == (self.n, self.d)
        with torch.cuda.device(self.device):
            check(lib().sm3_infonce_host(p1_host.data_ptr(), p2_host.data_ptr(), self.n, self.d,
                                         _lib._DTYPES[self.dtype], temperature, self.loss.data_ptr(),
                                         self.dp1.data_ptr(), self.dp2.data_ptr(), ptr(self.scratch),
                                         self.scratch.numel(), self.algo, stream_ptr()), "sm3_infonce_host")
        return self.loss, self.dp1, self.dp2


class HostInfoNCEPipeline:
    """Pipelined front end of the host-buffer entry (``sm3_host_pipe_*``): ``submit`` enqueues H2D -> fused fwd+bwd ->
    D2H for one batch on the handle's own three streams and returns a ticket; ``wait(ticket)`` blocks until that
    step's loss and gradients are in pinned host memory and returns them.  With ``depth`` >= 2 the copies of
    neighbouring steps overlap the kernels, so host-to-host throughput approaches the device-resident rate."""

    def __init__(self, n_pairs: int, d: int, dtype: torch.dtype = torch.bfloat16, algo: int = ALGO_AUTO, depth: int = 2,
                 device: Optional[torch.device] = None, group=None):
        """``group`` (torch.distributed process group, > 1 rank): peer mode -- ``n_pairs`` is this rank's share, every
        submit runs the multi-rank fused step (NVLink peer-memory exchange) between the copies; all ranks must submit the
        same number of steps in the same order."""
        import ctypes as C
        self.n, self.d, self.dtype, self.depth = n_pairs, d, dtype, depth
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        code = _lib._DTYPES[dtype]
        self.world, self.rank = _group_info(group)
        self.pbuf = None
        with torch.cuda.device(self.device):
            if self.world > 1:
                from . import peer
                self.pbuf = peer.get_peer_buffers(group, n_pairs * self.world, d, self.device)
                nbytes = lib().sm3_host_pipe_peer_scratch_bytes(n_pairs, n_pairs * self.world, d, code, depth)
            else:
                nbytes = lib().sm3_host_pipe_scratch_bytes(n_pairs, d, code, algo, depth)
            if nbytes == 0:
                raise ValueError("bad shape / depth for HostInfoNCEPipeline")
            self.scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            torch.cuda.synchronize(self.device)       # the handle's streams do not order against torch's allocator
            h = C.c_void_p()
            if self.world > 1:
                check(lib().sm3_host_pipe_create_peer(C.byref(h), n_pairs, n_pairs * self.world, d, code, depth,
                                                      ptr(self.scratch), self.scratch.numel()), "sm3_host_pipe_create_peer")
            else:
                check(lib().sm3_host_pipe_create(C.byref(h), n_pairs, d, code, algo, depth, ptr(self.scratch),
                                                 self.scratch.numel()), "sm3_host_pipe_create")
        self._h = h
        self._next = 0
        self.out = [(torch.empty(1, dtype=torch.float32).pin_memory(), torch.empty((n_pairs, d), dtype=dtype).pin_memory(),
                     torch.empty((n_pairs, d), dtype=dtype).pin_memory()) for _ in range(depth)]
        self.h2d_bytes = 2 * n_pairs * d * self.out[0][1].element_size()
        self.d2h_bytes = 4 + self.h2d_bytes

    def submit(self, p1_host: torch.Tensor, p2_host: torch.Tensor, temperature: float) -> int:
        assert not p1_host.is_cuda and p1_host.dtype == self.dtype and tuple(p1_host.shape) == (self.n, self.d)
        assert not p2_host.is_cuda and p2_host.dtype == self.dtype and tuple(p2_host.shape) == (self.n, self.d)
        with torch.cuda.device(self.device):
            loss, dp1, dp2 = self.out[self._next % self.depth]
            if self.pbuf is not None:
                pb = self.pbuf
                slot = pb.next_slot()
                fused = self.n % 128 == 0 and os.environ.get("SM3_PEER_FUSED", "1") != "0"
                mode = _fused_mode() if fused else 0
                t = check(lib().sm3_host_pipe_submit_peer(self._h, p1_host.data_ptr(), p2_host.data_ptr(), temperature,
                                                          loss.data_ptr(), dp1.data_ptr(), dp2.data_ptr(), self.rank,
                                                          self.world, ptr(pb.z[slot]), pb.zp[slot], ptr(pb.st[slot]),
                                                          pb.stp[slot], ptr(pb.flags), pb.fp, pb.step & 0x7FFFFFFF, mode),
                          "sm3_host_pipe_submit_peer")
            else:
                t = check(lib().sm3_host_pipe_submit(self._h, p1_host.data_ptr(), p2_host.data_ptr(), temperature,
                                                     loss.data_ptr(), dp1.data_ptr(), dp2.data_ptr()), "sm3_host_pipe_submit")
        self._next = t + 1
        return t

    def wait(self, ticket: int):
        """-> (loss[1] fp32, dp1, dp2) pinned host tensors of that ticket (valid until `depth` more submits)."""
        with torch.cuda.device(self.device):
            check(lib().sm3_host_pipe_wait(self._h, ticket), "sm3_host_pipe_wait")
        return self.out[ticket % self.depth]

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib().sm3_host_pipe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GraphedInfoNCE:
    """The single-GPU fused step (``sm3_infonce_step``: normalise -> K2 -> CE -> K3 -> normalise-backward) captured ONCE
    into a CUDA graph and replayed per step.  At the reference's batch sizes (cfg2: 4096 x 128 and below) the step is
    ~100 us of GPU work, less than what Python + autograd spend enqueueing it; a replay costs one launch.

    The graph reads the static buffers ``p1`` / ``p2`` (write your projector outputs into them, e.g. ``g.p1.copy_(x)``)
    and leaves ``loss`` (fp32 scalar, ``weight * mean CE``), ``dp1``, ``dp2`` (gradients w.r.t. p1 / p2) in static
    buffers that the next ``replay()`` overwrites."""

    def __init__(self, n_pairs: int, d: int, temperature: float, dtype: torch.dtype = torch.bfloat16,
                 precision: str = "auto", weight: float = 1.0, device: Optional[torch.device] = None):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.temperature, self.weight = float(temperature), float(weight)
        with torch.cuda.device(self.device):
            self.p1 = torch.zeros((n_pairs, d), dtype=dtype, device=self.device)
            self.p2 = torch.zeros((n_pairs, d), dtype=dtype, device=self.device)
            self.dp1, self.dp2 = torch.empty_like(self.p1), torch.empty_like(self.p2)
            self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
            _, self.algo = pick_precision(self.p1, precision)
            io = dtype_code(self.p1)
            nbytes = lib().sm3_infonce_step_scratch_bytes(n_pairs, d, io, self.algo)
            self._scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)

            def enqueue():
                check(lib().sm3_infonce_step(ptr(self.p1), ptr(self.p2), n_pairs, d, io, self.temperature, self.weight,
                                             ptr(self.loss), ptr(self.dp1), ptr(self.dp2), ptr(self._scratch),
                                             self._scratch.numel(), self.algo, stream_ptr()), "sm3_infonce_step")

            side = torch.cuda.Stream(device=self.device)          # warm-up off the default stream (capture rules)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                enqueue()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                enqueue()

    def replay(self):
        """-> (loss, dp1, dp2) static tensors, valid until the next replay()."""
        self.graph.replay()
        return self.loss, self.dp1, self.dp2

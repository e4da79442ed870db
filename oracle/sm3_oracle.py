"""CPU oracle for the SM3 contrastive hot path  --  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or the CPU baseline.  The product
(``skin_sm3_b200``) never imports this module and has no CPU fallback.

What it restates (all citations relative to the reference checkout):

* ``normalize``            -> ``F.normalize(features, dim=1)``        src/models/simclr.py:62,138,294
* ``infonce_logits``       -> the materialising logits builder         src/models/simclr.py:290-322
                              (= SimCLR.forward :54-93 with one projector)
* ``cross_entropy_col0``   -> ``nn.CrossEntropyLoss()(logits, 0)``     tools/backbone_train.py:531,101-121
* ``infonce_closed_form``  -> the same objective without the [M,M] matrix (chunked, any N) + gradients
* ``infonce_rowblock``     -> loss + gradient of a chosen set of rows at full benchmark size (cfg4)
* ``multihead_ce``         -> the 8-head loss loops                    tools/mlc_eval.py:159-162,
                                                                        tools/mlc_train.py:255-261,381
* ``bce_with_logits``      -> (no reference counterpart: PARITY UNPINNED; restates
                              torch.nn.functional.binary_cross_entropy_with_logits)
* ``knn_topk`` / ``knn_predict`` -> KNNOnlineEvaluator.predict         src/models/evaluator.py:43-83
* ``projector_tail`` (+``_bwd``) -> last Linear + BatchNorm1d(affine=False) of make_projector + F.normalize
                              src/models/simclr.py:25-26, :62, :294

Pinning: ``oracle/make_golden.py`` runs the *real* reference functions (imported from
/root/reference) on seeded inputs and stores inputs + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those files.
The arithmetic is numpy float64 unless stated.
"""
from __future__ import annotations

import numpy as np

NUM_CLASSES = (5, 3, 2, 3, 3, 3, 3, 2)  # tools/mlc_eval.py:63 (24 logits per sample)


# --------------------------------------------------------------------------------------
# A2: F.normalize(dim=1, eps=1e-12)
# --------------------------------------------------------------------------------------
def normalize(p: np.ndarray, eps: float = 1e-12):
    """z = p / max(||p||_2, eps) row-wise. Returns (z, inv_norm)."""
    p = np.asarray(p, dtype=np.float64)
    nrm = np.sqrt((p * p).sum(axis=1))
    inv = 1.0 / np.maximum(nrm, eps)
    return p * inv[:, None], inv


def normalize_bwd(dz: np.ndarray, z: np.ndarray, inv: np.ndarray, clamped=None):
    """dp = inv * (dz - z * <z, dz>)   (rows whose norm hit the eps clamp: dp = inv * dz)."""
    dz = np.asarray(dz, dtype=np.float64)
    dot = (z * dz).sum(axis=1, keepdims=True)
    dp = (dz - z * dot) * inv[:, None]
    if clamped is not None:
        dp = np.where(clamped[:, None], dz * inv[:, None], dp)
    return dp


# --------------------------------------------------------------------------------------
# N2: projector tail = last Linear (no bias) -> BatchNorm1d(affine=False) -> F.normalize   (simclr.py:25-26 then :62/:294)
# --------------------------------------------------------------------------------------
def projector_tail(h: np.ndarray, w: np.ndarray, bn_eps: float = 1e-5, l2_eps: float = 1e-12, running=None,
                   momentum: float = 0.1, training: bool = True):
    """z = normalize(batchnorm(h @ w.T)) in float64.  Training mode uses the batch statistics (biased variance) and returns
    the updated running statistics (unbiased variance, momentum) as nn.BatchNorm1d does; eval mode uses `running`.
    Returns dict(y, mean, rstd, yhat, z, inv, running_mean, running_var)."""
    h = np.asarray(h, np.float64)
    w = np.asarray(w, np.float64)
    y = h @ w.T
    r = y.shape[0]
    if training:
        mean = y.mean(axis=0)
        var = y.var(axis=0)
        rm = rv = None
        if running is not None:
            rm = (1 - momentum) * np.asarray(running[0], np.float64) + momentum * mean
            rv = (1 - momentum) * np.asarray(running[1], np.float64) + momentum * var * r / max(r - 1, 1)
    else:
        mean, var = np.asarray(running[0], np.float64), np.asarray(running[1], np.float64)
        rm, rv = mean, var
    rstd = 1.0 / np.sqrt(var + bn_eps)
    yhat = (y - mean) * rstd
    z, inv = normalize(yhat, l2_eps)
    return dict(y=y, mean=mean, rstd=rstd, yhat=yhat, z=z, inv=inv, running_mean=rm, running_var=rv)


def projector_tail_bwd(h, w, fwd: dict, dz: np.ndarray, training: bool = True, l2_eps: float = 1e-12):
    """Gradients of sum(z * dz) w.r.t. h and w through normalize -> batchnorm -> linear (float64)."""
    h = np.asarray(h, np.float64)
    w = np.asarray(w, np.float64)
    clamped = (1.0 / fwd["inv"]) <= l2_eps
    dyhat = normalize_bwd(dz, fwd["z"], fwd["inv"], clamped)
    if training:
        dy = fwd["rstd"] * (dyhat - dyhat.mean(axis=0) - fwd["yhat"] * (dyhat * fwd["yhat"]).mean(axis=0))
    else:
        dy = fwd["rstd"] * dyhat
    return dy @ w, dy.T @ h, dy


# --------------------------------------------------------------------------------------
# A3 + A4: literal materialising logits builder (small M only)
# --------------------------------------------------------------------------------------
def infonce_logits(p1: np.ndarray, p2: np.ndarray, temperature: float):
    """Restates SimCLRSkinV3._cal_logits after the projectors (simclr.py:293-320).

    Returns logits [M, M-1] with the positive in column 0 followed by the negatives in
    ascending column order, and the all-zero labels [M].
    """
    p = np.concatenate([np.asarray(p1, np.float64), np.asarray(p2, np.float64)], axis=0)
    n = p1.shape[0]
    m = 2 * n
    z, _ = normalize(p)
    sim = z @ z.T                                         # simclr.py:300
    ids = np.concatenate([np.arange(n), np.arange(n)])    # simclr.py:296
    same = ids[:, None] == ids[None, :]                   # simclr.py:297
    off = ~np.eye(m, dtype=bool)                          # simclr.py:303
    same_od = same[off].reshape(m, m - 1)                 # simclr.py:304
    sim_od = sim[off].reshape(m, m - 1)                   # simclr.py:305-307
    pos = sim_od[same_od].reshape(m, -1)                  # simclr.py:310
    neg = sim_od[~same_od].reshape(m, -1)                 # simclr.py:313-315
    logits = np.concatenate([pos, neg], axis=1) / temperature  # simclr.py:317,320
    labels = np.zeros(m, dtype=np.int64)                  # simclr.py:318
    return logits, labels


def cross_entropy_col0(logits: np.ndarray) -> float:
    """nn.CrossEntropyLoss()(logits, zeros): mean_i(logsumexp_j logits_ij - logits_i0)."""
    mx = logits.max(axis=1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(logits - mx).sum(axis=1))
    return float((lse - logits[:, 0]).mean())


# --------------------------------------------------------------------------------------
# closed form (never builds [M, M]); supports a row slice for the multi-GPU shard test
# --------------------------------------------------------------------------------------
def global_row_index(n_local: int, pair_offset: int, n_pairs_global: int) -> np.ndarray:
    """Global row ids of a rank's 2*n_local rows: [first halves ; second halves]."""
    a = pair_offset + np.arange(n_local)
    return np.concatenate([a, n_pairs_global + a])


def infonce_stats(z: np.ndarray, n_pairs: int, temperature: float, rows=None, chunk: int = 1024):
    """Per-row sufficient statistics of the InfoNCE term on normalised rows z [M, D].

    pos_i     = S_{i,pos(i)}                         (already divided by T)
    lse_neg_i = log sum_{j not in {i, pos(i)}} exp(S_ij)
    with S = z z^T / T, pos(i) = (i + n_pairs) mod M.  `rows` = global row ids to evaluate.
    CE([pos, lse_neg], 0) == CE(reference logits, 0) exactly.
    """
    z = np.asarray(z, np.float64)
    m = z.shape[0]
    assert m == 2 * n_pairs
    rows = np.arange(m) if rows is None else np.asarray(rows)
    pos = np.empty(len(rows))
    lse = np.empty(len(rows))
    for s in range(0, len(rows), chunk):
        r = rows[s:s + chunk]
        S = (z[r] @ z.T) / temperature
        pj = (r + n_pairs) % m
        ar = np.arange(len(r))
        pos[s:s + chunk] = S[ar, pj]
        S[ar, r] = -np.inf
        S[ar, pj] = -np.inf
        mx = S.max(axis=1)
        mx = np.where(np.isfinite(mx), mx, 0.0)
        with np.errstate(divide="ignore"):
            lse[s:s + chunk] = mx + np.log(np.exp(S - mx[:, None]).sum(axis=1))
    return pos, lse


def stats_to_loss(pos: np.ndarray, lse_neg: np.ndarray) -> float:
    """mean_i( log(e^pos + e^lse_neg) - pos )  ==  stock CE on the [M,2] logits."""
    return float(np.mean(np.logaddexp(pos, lse_neg) - pos))


def infonce_closed_form(p1, p2, temperature: float, chunk: int = 1024, upstream: float = 1.0):
    """Loss and d loss / d p1, d p2 of one InfoNCE term (normalise included), float64.

    dL/dS = (P - Y)/M ; dL/dz = (G + G^T) z / T ; then the normalise backward.
    """
    p1 = np.asarray(p1, np.float64)
    p2 = np.asarray(p2, np.float64)
    n = p1.shape[0]
    m = 2 * n
    p = np.concatenate([p1, p2], axis=0)
    z, inv = normalize(p)
    pos, lse_neg = infonce_stats(z, n, temperature, chunk=chunk)
    lse = np.logaddexp(pos, lse_neg)
    loss = float(np.mean(lse - pos))
    dz = np.zeros_like(z)
    idx = np.arange(m)
    pj_all = (idx + n) % m
    for s in range(0, m, chunk):
        r = idx[s:s + chunk]
        S = (z[r] @ z.T) / temperature
        ar = np.arange(len(r))
        # G_ij + G_ji with G = (P - Y)/M ; S symmetric  =>  P_ij = exp(S_ij - lse_i), P_ji = exp(S_ij - lse_j)
        H = np.exp(S - lse[r][:, None]) + np.exp(S - lse[None, :])
        H[ar, r] = 0.0
        H[ar, pj_all[r]] -= 2.0
        dz[r] = (H @ z) * (upstream / (m * temperature))
    clamped = (1.0 / inv) <= 1e-12
    dp = normalize_bwd(dz, z, inv, clamped)
    return loss, dp[:n], dp[n:]


def infonce_rowblock(p1, p2, temperature: float, rows, chunk: int = 2048, matmul_dtype=np.float64,
                     upstream: float = 1.0, threads=None):
    """Loss of one InfoNCE term and d loss / d p for a chosen set of GLOBAL rows, at sizes where the full fp64
    closed form above is too slow (cfg4: M = 65536 rows, 4.3e9 logits).

    ``rows`` index the concatenated batch ``p = cat([p1, p2])`` (0 <= row < 2N).  Returns
    ``(loss, dp_rows [len(rows), D], lse_all [M])`` with the same maths as ``infonce_closed_form``
    (simclr.py:290-322 + CE backbone_train.py:531 + their autograd backward):

        lse_j  = log sum_{k != j} exp(S_jk)                  for ALL j (one full pass, needed by every row's gradient)
        dz_i   = sum_j (exp(S_ij - lse_i) + exp(S_ij - lse_j) - 2 [j = pos(i)]) z_j * upstream / (M T)     (j != i)
        dp_i   = (dz_i - z_i <z_i, dz_i>) / max(||p_i||, 1e-12)

    ``matmul_dtype=np.float64`` is the pinned mode (agrees with the reference goldens to 1e-12, see
    tests/test_oracle_golden.py).  ``np.float32`` runs the M x M similarity pass as a multi-threaded fp32 GEMM with
    fp64 row-sum accumulation (torch CPU): relative error <= 1e-5 on loss and gradients (pinned by the same test),
    three orders of magnitude below the 2e-2 bf16 tolerance it is used to check at full size.
    """
    import torch
    if threads:
        torch.set_num_threads(int(threads))
    p = np.concatenate([np.asarray(p1, np.float64), np.asarray(p2, np.float64)], axis=0)
    n = p1.shape[0]
    m = 2 * n
    z, inv = normalize(p)
    tdt = torch.float64 if matmul_dtype == np.float64 else torch.float32
    zt = torch.from_numpy(np.ascontiguousarray(z)).to(tdt)
    inv_t = 1.0 / temperature
    lse = np.empty(m)
    pos = np.empty(m)
    for s in range(0, m, chunk):
        e = min(m, s + chunk)
        S = (zt[s:e] @ zt.T) * inv_t                     # simclr.py:300, :320
        ar = torch.arange(e - s)
        pj = (torch.arange(s, e) + n) % m
        pos[s:e] = S[ar, pj].double().numpy()
        S[ar, torch.arange(s, e)] = -float("inf")        # drop the diagonal (simclr.py:303-307)
        mx = S.max(dim=1, keepdim=True).values
        mx = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
        lse[s:e] = (mx[:, 0].double() + torch.exp(S - mx).sum(dim=1, dtype=torch.float64).log()).numpy()
    loss = float(np.mean(lse - pos))
    rows = np.asarray(rows, np.int64)
    dp = np.empty((len(rows), p.shape[1]))
    lse_t = torch.from_numpy(lse)
    z64 = torch.from_numpy(np.ascontiguousarray(z))
    for s in range(0, len(rows), chunk):
        r = rows[s:s + chunk]
        rt = torch.from_numpy(r)
        S = ((zt[rt] @ zt.T) * inv_t).double()
        H = torch.exp(S - lse_t[rt][:, None]) + torch.exp(S - lse_t[None, :])
        ar = torch.arange(len(r))
        H[ar, rt] = 0.0
        H[ar, (rt + n) % m] -= 2.0
        dz = (H @ z64).numpy() * (upstream / (m * temperature))
        clamped = (1.0 / inv[r]) <= 1e-12
        dp[s:s + chunk] = normalize_bwd(dz, z[r], inv[r], clamped)
    return loss, dp, lse


def stats_backward(z, n_pairs, temperature, g_pos, g_lse, rows=None, chunk: int = 1024):
    """d/dz of sum_i (g_pos_i * pos_i + g_lse_i * lse_neg_i) for the FULL (all rows) problem.

    This is what the sufficient-statistics autograd node must return for upstream grads
    (g_pos, g_lse) handed back by the stock cross-entropy.
    """
    z = np.asarray(z, np.float64)
    m = z.shape[0]
    pos, lse_neg = infonce_stats(z, n_pairs, temperature, chunk=chunk)
    idx = np.arange(m)
    pj = (idx + n_pairs) % m
    dz = np.zeros_like(z)
    for s in range(0, m, chunk):
        r = idx[s:s + chunk]
        S = (z[r] @ z.T) / temperature
        ar = np.arange(len(r))
        with np.errstate(over="ignore", invalid="ignore"):
            Gr = g_lse[r][:, None] * np.exp(S - lse_neg[r][:, None])       # G_ij
            Gc = g_lse[None, :] * np.exp(S - lse_neg[None, :])             # G_ji
        Gr = np.nan_to_num(Gr, nan=0.0, posinf=0.0)
        Gc = np.nan_to_num(Gc, nan=0.0, posinf=0.0)
        H = Gr + Gc
        H[ar, r] = 0.0
        H[ar, pj[r]] = g_pos[r] + g_pos[pj[r]]
        dz[r] = (H @ z) / temperature
    return dz


# --------------------------------------------------------------------------------------
# H1: multi-head softmax cross-entropy
# --------------------------------------------------------------------------------------
def proto_heads(sa_feats: np.ndarray, w_cat: np.ndarray, class_counts, l2_norm: bool, eps: float = 1e-12):
    """Tail of the multi-label Model.forward (reference tools/mlc_train.py:81-87): optional row normalisation of every
    feature slot, then ``preds[h] = sa[h % Hf] @ W_h^T`` for the bias-free prototype Linears; logits concatenated [B, C]."""
    sa = np.asarray(sa_feats, np.float64)
    hf = sa.shape[0]
    if l2_norm:
        sa = sa / np.maximum(np.linalg.norm(sa, axis=-1, keepdims=True), eps)
    outs, off = [], 0
    for h, n in enumerate(class_counts):
        outs.append(sa[h % hf] @ np.asarray(w_cat[off:off + n], np.float64).T)
        off += n
    return sa, np.concatenate(outs, axis=1)


def proto_heads_bwd(sa_feats, w_cat, class_counts, l2_norm: bool, dlogits, eps: float = 1e-12):
    """Gradients of ``proto_heads`` w.r.t. the (un-normalised) features and the concatenated prototype weights."""
    x = np.asarray(sa_feats, np.float64)
    hf = x.shape[0]
    nrm = np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), eps)
    z = x / nrm if l2_norm else x
    dz = np.zeros_like(x)
    dw = np.zeros_like(np.asarray(w_cat, np.float64))
    off = 0
    for h, n in enumerate(class_counts):
        g = np.asarray(dlogits[:, off:off + n], np.float64)
        dz[h % hf] += g @ np.asarray(w_cat[off:off + n], np.float64)
        dw[off:off + n] = g.T @ z[h % hf]
        off += n
    dx = (dz - z * (z * dz).sum(-1, keepdims=True)) / nrm if l2_norm else dz
    return dx, dw


def multihead_ce(logits: np.ndarray, labels: np.ndarray, weights=None, inv_temperature: float = 1.0,
                 ignore_index: int = -100, num_classes=NUM_CLASSES):
    """loss = sum_h w_h * CE(logits_h * inv_T, labels[:, h]) / H  with per-head mean reduction.

    logits: [B, sum(num_classes)] heads concatenated along dim 1; labels: [B, H] int64.
    Per head CE is the mean over rows whose label != ignore_index
    (tools/mlc_eval.py:159-162 with label_weights; tools/mlc_train.py:255-261 with pred/T,
    ignore_index=-100 :381).  Returns (loss, dlogits).
    """
    logits = np.asarray(logits, np.float64)
    labels = np.asarray(labels)
    B = logits.shape[0]
    H = len(num_classes)
    w = np.ones(H) if weights is None else np.asarray(weights, np.float64)
    loss = 0.0
    grad = np.zeros_like(logits)
    o = 0
    for h, c in enumerate(num_classes):
        x = logits[:, o:o + c] * inv_temperature
        y = labels[:, h]
        valid = y != ignore_index
        cnt = int(valid.sum())
        mx = x.max(axis=1, keepdims=True)
        e = np.exp(x - mx)
        sm = e / e.sum(axis=1, keepdims=True)
        lse = mx[:, 0] + np.log(e.sum(axis=1))
        ys = np.where(valid, y, 0)
        nll = lse - x[np.arange(B), ys]
        if cnt > 0:
            loss += w[h] * float(nll[valid].sum()) / cnt / H
            g = sm.copy()
            g[np.arange(B), ys] -= 1.0
            g *= valid[:, None]
            grad[:, o:o + c] = g * (w[h] * inv_temperature / cnt / H)
        else:
            loss += float("nan")  # torch: mean over zero valid rows -> nan
        o += c
    return loss, grad


# --------------------------------------------------------------------------------------
# H2: BCE-with-logits (PARITY UNPINNED by the reference: it has no BCE anywhere)
# --------------------------------------------------------------------------------------
def bce_with_logits(x: np.ndarray, t: np.ndarray, pos_weight=None):
    """mean over all elements of  max(x,0) - x*t + log1p(exp(-|x|))  (+ pos_weight form).

    Returns (loss, dx).  Mirrors torch.nn.functional.binary_cross_entropy_with_logits(reduction='mean').
    """
    x = np.asarray(x, np.float64)
    t = np.asarray(t, np.float64)
    sp_neg = np.maximum(-x, 0) + np.log1p(np.exp(-np.abs(x)))      # softplus(-x)
    sig = 1.0 / (1.0 + np.exp(-x))
    if pos_weight is None:
        el = (1 - t) * x + sp_neg
        dx = sig - t
    else:
        pw = np.asarray(pos_weight, np.float64)[None, :]
        lw = 1 + (pw - 1) * t
        el = (1 - t) * x + lw * sp_neg
        dx = (1 - t) - lw * (1 - sig)
    return float(el.mean()), dx / x.size


# --------------------------------------------------------------------------------------
# N1: KNN retrieval (evaluator.py:43-83)
# --------------------------------------------------------------------------------------
def knn_topk(query: np.ndarray, bank: np.ndarray, k: int):
    """sim = query @ bank.T ; top-k along the bank axis (values descending, ties -> lower index)."""
    sim = np.asarray(query, np.float64) @ np.asarray(bank, np.float64).T
    order = np.argsort(-sim, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(sim, order, axis=1), order


def knn_predict(query, bank, bank_labels, num_classes: int, k: int, temperature: float):
    """Weighted k-NN vote; returns class ids sorted by descending score  (evaluator.py:60-83)."""
    w, idx = knn_topk(query, bank, k)
    lab = np.asarray(bank_labels)[idx]
    w = np.exp(w / temperature)
    B = query.shape[0]
    scores = np.zeros((B, num_classes))
    for c in range(num_classes):
        scores[:, c] = (w * (lab == c)).sum(axis=1)
    return np.argsort(-scores, axis=1, kind="stable"), scores


# --------------------------------------------------------------------------------------
# N4: DeepCluster spherical k-means of the memory bank (tools/mlc_train.py:116-189)
# --------------------------------------------------------------------------------------
def spherical_kmeans(emb: np.ndarray, init_idx: np.ndarray, n_iters: int = 10, dtype=np.float32):
    """cluster_memory's rank-0 loop restated: centroids = emb[init_idx] (mlc_train.py:144-145); n_iters times
    E step = argmax of emb @ centroids.T (:150-152), M step = mean of each non-empty cluster (:157-172) followed by
    L2 normalisation of ALL centroids (:175); one final E step (:150-155).  Returns (assignments [n], centroids [K, D]).
    dtype float32 mirrors the reference's arithmetic; argmax ties resolve to the lower centroid index."""
    emb = np.asarray(emb, dtype)
    cent = emb[np.asarray(init_idx)].copy()
    k = cent.shape[0]
    assign = None
    for it in range(n_iters + 1):
        assign = np.argmax(emb @ cent.T, axis=1)
        if it == n_iters:
            break
        for c in range(k):
            members = emb[assign == c]
            if len(members) > 0:
                cent[c] = members.sum(axis=0, dtype=dtype) / dtype(len(members))
        norm = np.sqrt((cent.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(dtype)
        cent = cent / np.maximum(norm, dtype(1e-12))
    return assign.astype(np.int64), cent


def cluster_memory(index: np.ndarray, emb: np.ndarray, init_idx: np.ndarray, n_iters: int = 10):
    """assignments[index[i]] = cluster of emb[i] (mlc_train.py:119-121, 178-182); -100 where no sample points."""
    a, cent = spherical_kmeans(emb, init_idx, n_iters)
    out = np.full(len(index), -100, np.int64)
    out[np.asarray(index)] = a
    return out, cent

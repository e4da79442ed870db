"""Pins oracle/ (the CPU restatement) to outputs of the REAL reference stored in tests/golden/.

CPU only.  These goldens were produced by oracle/make_golden.py importing /root/reference.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port, sm3_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL = ["infonce_n4_d8_T05", "infonce_n64_d128_T01", "infonce_n48_d128_T01_corr"]
BIG = ["infonce_n200_d64_T02_corr", "infonce_n512_d256_T01"]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("name", FULL)
def test_literal_logits_match_reference(name):
    g = load(name)
    logits, labels = O.infonce_logits(g["p1"], g["p2"], float(g["temperature"]))
    assert logits.shape == g["logits_f64"].shape
    np.testing.assert_allclose(logits, g["logits_f64"], rtol=0, atol=1e-12)
    assert labels.dtype == np.int64 and not labels.any()
    assert abs(O.cross_entropy_col0(logits) - float(g["loss_f64"])) < 1e-12


@pytest.mark.parametrize("name", FULL + BIG)
def test_closed_form_loss_and_grads(name):
    g = load(name)
    T = float(g["temperature"])
    loss, dp1, dp2 = O.infonce_closed_form(g["p1"], g["p2"], T, chunk=96)
    assert abs(loss - float(g["loss_f64"])) < 1e-12 * max(1.0, abs(loss))
    if "dp1_f64" in g:
        np.testing.assert_allclose(dp1, g["dp1_f64"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(dp2, g["dp2_f64"], rtol=0, atol=1e-14)
    else:
        r = g["grad_rows"]
        np.testing.assert_allclose(dp1[r], g["dp1_rows_f64"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(dp2[r], g["dp2_rows_f64"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(dp1.sum(0), g["dp1_sum_f64"], rtol=0, atol=1e-12)
        assert abs(np.abs(dp1).sum() + np.abs(dp2).sum() - float(g["dp_abs_sum_f64"])) < 1e-10


@pytest.mark.parametrize("name", FULL + BIG)
def test_sufficient_statistics_equal_reference_ce(name):
    """CE([pos, lse_neg], 0) == CE(reference logits [M, M-1], 0); column 0 == reference positives."""
    g = load(name)
    T = float(g["temperature"])
    n = int(g["n"])
    z, _ = O.normalize(np.concatenate([g["p1"], g["p2"]]).astype(np.float64))
    pos, lse_neg = O.infonce_stats(z, n, T, chunk=100)
    assert abs(O.stats_to_loss(pos, lse_neg) - float(g["loss_f64"])) < 1e-12 * max(1, float(g["loss_f64"]))
    col0 = g["logits_f64"][:, 0] if "logits_f64" in g else g["logits_col0_f64"]
    np.testing.assert_allclose(pos, col0, rtol=0, atol=1e-12)
    # negatives in ascending column order (row 0: columns 1..M-1 without n)
    row0 = g["logits_f64"][0] if "logits_f64" in g else g["logits_row0_f64"]
    s0 = (z[0] @ z.T) / T
    expect = np.concatenate([[s0[n]], np.delete(s0, [0, n])])
    np.testing.assert_allclose(row0, expect, rtol=0, atol=1e-12)


def test_stats_backward_matches_autograd_through_ce():
    g = load("infonce_n48_d128_T01_corr")
    T, n = float(g["temperature"]), int(g["n"])
    p = np.concatenate([g["p1"], g["p2"]]).astype(np.float64)
    z, inv = O.normalize(p)
    pos, lse_neg = O.infonce_stats(z, n, T)
    # upstream grads the stock CE hands back for logits [pos, lse_neg], target 0, mean reduction
    m = 2 * n
    sig = 1.0 / (1.0 + np.exp(pos - lse_neg))          # softmax prob of column 1
    g_lse, g_pos = sig / m, -sig / m
    dz = O.stats_backward(z, n, T, g_pos, g_lse)
    dp = O.normalize_bwd(dz, z, inv)
    np.testing.assert_allclose(dp[:n], g["dp1_f64"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(dp[n:], g["dp2_f64"], rtol=0, atol=1e-14)


def test_edge_cases():
    g = load("infonce_edge")
    for case in g["case_names"]:
        for T in (0.1, 0.5):
            k = f"{case}_T{T}"
            p1, p2 = g[k + "_p1"], g[k + "_p2"]
            logits, _ = O.infonce_logits(p1, p2, T)
            np.testing.assert_allclose(logits, g[k + "_logits"], rtol=0, atol=1e-12)
            loss, dp1, dp2 = O.infonce_closed_form(p1, p2, T)
            assert abs(loss - float(g[k + "_loss"])) < 1e-12, k
            # the zero row hits F.normalize's eps clamp: its gradient is dz / 1e-12 (~1e12), so rtol
            np.testing.assert_allclose(dp1, g[k + "_dp1"], rtol=1e-12, atol=1e-12, err_msg=k)
            np.testing.assert_allclose(dp2, g[k + "_dp2"], rtol=1e-12, atol=1e-12, err_msg=k)


def test_analytic_known_answers():
    """SURVEY 8c KATs: identical rows / orthonormal rows -> log(M-1); duplicated pairs."""
    n, d, T = 8, 32, 0.1
    m = 2 * n
    same = np.ones((n, d))
    loss, _, _ = O.infonce_closed_form(same, same, T)
    assert abs(loss - np.log(m - 1)) < 1e-12
    eye = np.eye(m, d)
    loss, _, _ = O.infonce_closed_form(eye[:n], eye[n:], T)
    assert abs(loss - np.log(m - 1)) < 1e-12
    e = np.eye(n, d)
    loss, _, _ = O.infonce_closed_form(e, e, T)
    assert abs(loss - (np.log(np.exp(1 / T) + m - 2) - 1 / T)) < 1e-12


def test_sharded_rows_reproduce_global(tmp_path):
    """W-rank row-block sharding with global column order [all f1 ; all f2]."""
    g = load("infonce_n64_d128_T01")
    T, n = float(g["temperature"]), int(g["n"])
    z, _ = O.normalize(np.concatenate([g["p1"], g["p2"]]).astype(np.float64))
    W = 4
    nl = n // W
    losses = []
    for r in range(W):
        rows = O.global_row_index(nl, r * nl, n)
        pos, lse = O.infonce_stats(z, n, T, rows=rows)
        losses.append(O.stats_to_loss(pos, lse))
    assert abs(np.mean(losses) - float(g["loss_f64"])) < 1e-12


def test_heads_match_reference_loops():
    g = load("heads")
    loss, grad = O.multihead_ce(g["eval_logits"], g["eval_labels"], g["eval_weights"])
    assert abs(loss - float(g["eval_loss"])) < 1e-12
    np.testing.assert_allclose(grad, g["eval_grad"], rtol=0, atol=1e-14)
    loss, grad = O.multihead_ce(g["dc_logits"], g["dc_labels"], None, 1.0 / float(g["dc_temperature"]), -100)
    assert abs(loss - float(g["dc_loss"])) < 1e-11
    np.testing.assert_allclose(grad, g["dc_grad"], rtol=0, atol=1e-13)


def test_bce_matches_torch():
    """H2 has no reference counterpart (parity unpinned): oracle == torch BCE-with-logits."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(33, 24)) * 3
    t = (rng.random((33, 24)) < 0.3).astype(np.float64)
    xt = torch.tensor(x, requires_grad=True)
    l = torch.nn.functional.binary_cross_entropy_with_logits(xt, torch.tensor(t))
    l.backward()
    loss, dx = O.bce_with_logits(x, t)
    assert abs(loss - l.item()) < 1e-12
    np.testing.assert_allclose(dx, xt.grad.numpy(), rtol=0, atol=1e-14)
    pw = rng.random(24) * 3 + 0.5
    xt = torch.tensor(x, requires_grad=True)
    l = torch.nn.functional.binary_cross_entropy_with_logits(xt, torch.tensor(t), pos_weight=torch.tensor(pw))
    l.backward()
    loss, dx = O.bce_with_logits(x, t, pw)
    assert abs(loss - l.item()) < 1e-12
    np.testing.assert_allclose(dx, xt.grad.numpy(), rtol=0, atol=1e-14)


def test_knn_matches_reference_evaluator():
    g = load("knn")
    val, idx = O.knn_topk(g["query"], g["bank"], int(g["k"]))
    assert (idx == g["topk_idx"]).all()
    np.testing.assert_allclose(val, g["topk_val"], rtol=0, atol=1e-14)
    pred, _ = O.knn_predict(g["query"], g["bank"], g["bank_labels"], int(g["num_classes"]), int(g["k"]),
                            float(g["temperature"]))
    assert (pred == g["pred_labels"]).all()


@pytest.mark.parametrize("name", FULL)
def test_torch_port_matches_reference(name):
    """oracle/ref_port.py (the CPU baseline bench.py times) == the real reference, fp32 and fp64."""
    g = load(name)
    T = float(g["temperature"])
    for tag, dt, tol in (("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-13)):
        p1 = torch.from_numpy(g["p1"]).to(dt)
        p2 = torch.from_numpy(g["p2"]).to(dt)
        logits, target = ref_port.port_cal_logits(p1, p2, T)
        np.testing.assert_allclose(logits.numpy(), g[f"logits_{tag}"], rtol=0, atol=tol * 10)
        loss, d1, d2 = ref_port.port_infonce_step(p1, p2, T)
        assert abs(loss.item() - float(g[f"loss_{tag}"])) <= tol * max(1, abs(loss.item()))
        np.testing.assert_allclose(d1.numpy(), g[f"dp1_{tag}"], rtol=0, atol=tol)
    # the row-block sample (used for shapes the reference cannot hold) sums to the full step
    p1 = torch.from_numpy(g["p1"]).double(); p2 = torch.from_numpy(g["p2"]).double()
    m = 2 * p1.shape[0]
    tot, a1 = 0.0, 0
    blk = max(1, m // 4)
    for s in range(0, m, blk):
        l, d1, _ = ref_port.port_infonce_step_rowblock(p1, p2, T, s, min(blk, m - s))
        tot += l.item(); a1 = a1 + d1
    assert abs(tot - float(g["loss_f64"])) < 1e-12
    np.testing.assert_allclose(a1.numpy(), g["dp1_f64"], rtol=0, atol=1e-13)


def test_torch_port_heads():
    g = load("heads")
    nc = list(g["num_classes"])
    outs = list(torch.split(torch.from_numpy(g["eval_logits"]), nc, dim=1))
    l = ref_port.port_multihead_ce(outs, torch.from_numpy(g["eval_labels"]), g["eval_weights"])
    assert abs(l.item() - float(g["eval_loss"])) < 1e-12


def test_kmeans_oracle_matches_reference_cluster_memory():
    """oracle.cluster_memory vs tools/mlc_train.py::cluster_memory (real function, run by oracle/make_golden.py):
    assignments exact -- including the case whose clusters run empty -- and centroids to fp32 rounding."""
    g = np.load(os.path.join(GOLDEN, "kmeans.npz"))
    for c in g["cases"]:
        a, cent = O.cluster_memory(g[f"{c}_index"], g[f"{c}_emb"], g[f"{c}_init_idx"])
        assert (a == g[f"{c}_assign"]).all(), c
        assert np.abs(cent - g[f"{c}_centroids"]).max() < 1e-6, c
        # the initial indices are what torch.randperm gives for the recorded seed (how the product reproduces them)
        torch.manual_seed(int(g[f"{c}_seed"]))
        assert (torch.randperm(len(g[f"{c}_emb"]))[: len(g[f"{c}_init_idx"])].numpy() == g[f"{c}_init_idx"]).all()


@pytest.mark.parametrize("name", ["infonce_n48_d128_T01_corr", "infonce_n200_d64_T02_corr", "infonce_n512_d256_T01"])
def test_rowblock_oracle_pinned_to_reference(name):
    """oracle.infonce_rowblock (the full-size checker used at cfg4 / cfg2 and by bench.py's parity block) against the
    reference goldens: fp64 mode to 1e-12, fp32-GEMM mode (what runs at M = 65536) to 1e-6 (loss) / 1e-5 (gradients)."""
    g = load(name)
    T, n = float(g["temperature"]), int(g["n"])
    if "dp1_f64" in g:
        r = np.arange(n)
        ref = np.concatenate([g["dp1_f64"], g["dp2_f64"]])
    else:
        r = g["grad_rows"]
        ref = np.concatenate([g["dp1_rows_f64"], g["dp2_rows_f64"]])
    rows = np.concatenate([r, n + r])
    loss, dp, lse = O.infonce_rowblock(g["p1"], g["p2"], T, rows, chunk=77)
    assert abs(loss - float(g["loss_f64"])) < 1e-12 * max(1.0, abs(loss))
    np.testing.assert_allclose(dp, ref, rtol=0, atol=1e-14)
    loss32, dp32, _ = O.infonce_rowblock(g["p1"], g["p2"], T, rows, chunk=77, matmul_dtype=np.float32, upstream=3.0)
    assert abs(loss32 - float(g["loss_f64"])) < 1e-6 * abs(loss32)
    assert np.abs(dp32 / 3.0 - ref).max() < 1e-5 * np.abs(ref).max()


def test_projector_tail_oracle_pinned_to_reference():
    """oracle.projector_tail / projector_tail_bwd against the REAL make_projector's last two layers (train and eval mode)
    + F.normalize, values and autograd gradients recorded in fp64 (tests/golden/projtail.npz)."""
    g = load("projtail")
    for c in g["cases"]:
        h, w, gz = g[f"{c}_h"].astype(np.float64), g[f"{c}_w"].astype(np.float64), g[f"{c}_gz"]
        f = O.projector_tail(h, w, running=(g[f"{c}_rm0"], g[f"{c}_rv0"]))
        np.testing.assert_allclose(f["z"], g[f"{c}_z"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(f["running_mean"], g[f"{c}_rm1"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(f["running_var"], g[f"{c}_rv1"], rtol=0, atol=1e-13)
        dh, dw, _ = O.projector_tail_bwd(h, w, f, gz)
        np.testing.assert_allclose(dh, g[f"{c}_dh"], rtol=0, atol=1e-11)
        np.testing.assert_allclose(dw, g[f"{c}_dw"], rtol=0, atol=1e-11)
        fe = O.projector_tail(h, w, running=(g[f"{c}_rm1"], g[f"{c}_rv1"]), training=False)
        np.testing.assert_allclose(fe["z"], g[f"{c}_z_eval"], rtol=0, atol=1e-12)
    # the whole term: two tails -> _cal_logits -> CE, gradients down to the tails' inputs and weights
    n, T = int(g["term_n"]), float(g["term_T"])
    f1, f2 = g["term_f1"].astype(np.float64), g["term_f2"].astype(np.float64)
    w1, w2 = g["term_w1"].astype(np.float64), g["term_w2"].astype(np.float64)
    t1, t2 = O.projector_tail(f1, w1), O.projector_tail(f2, w2)
    loss, dp1, dp2 = O.infonce_closed_form(t1["yhat"], t2["yhat"], T)
    assert abs(loss - float(g["term_loss"])) < 1e-12
    # infonce_closed_form returns d loss / d (un-normalised rows); push it through the BatchNorm + Linear backward
    for t, dp, f, w, df, dw in ((t1, dp1, f1, w1, "term_df1", "term_dw1"), (t2, dp2, f2, w2, "term_df2", "term_dw2")):
        dy = t["rstd"] * (dp - dp.mean(axis=0) - t["yhat"] * (dp * t["yhat"]).mean(axis=0))
        np.testing.assert_allclose(dy @ w, g[df], rtol=0, atol=1e-12)
        np.testing.assert_allclose(dy.T @ f, g[dw], rtol=0, atol=1e-12)


def test_proto_heads_oracle_pinned_to_reference_model():
    """oracle.proto_heads / proto_heads_bwd against the REAL multi-label Model (tools/mlc_train.py:58-89) + the DeepCluster
    loss loop (:255-261): returned features, predictions, loss through the 8-head CE oracle, and both gradients."""
    g = np.load(os.path.join(GOLDEN, "mlchead.npz"), allow_pickle=False)
    counts = [5, 3, 2, 3, 3, 3, 3, 2]
    for tag in g["cases"]:
        sa_in, w, l2 = g[f"{tag}_sa_in"], g[f"{tag}_w"], bool(g[f"{tag}_l2"])
        sa_out, logits = O.proto_heads(sa_in, w, counts, l2)
        assert np.abs(sa_out - g[f"{tag}_sa_out"]).max() < 2e-6, tag          # fixture stores the features in fp32
        assert np.abs(logits - g[f"{tag}_logits"]).max() < 2e-6, tag
        loss, dlogits = O.multihead_ce(logits, g[f"{tag}_targets"], None, 1.0 / float(g[f"{tag}_T"]), ignore_index=-100)
        assert abs(loss - float(g[f"{tag}_loss"])) < 1e-6, tag
        dx, dw = O.proto_heads_bwd(sa_in, w, counts, l2, dlogits)
        assert np.abs(dx - g[f"{tag}_d_sa_in"]).max() < 1e-6 * max(1.0, np.abs(g[f"{tag}_d_sa_in"]).max() * 1e3), tag
        assert np.abs(dw - g[f"{tag}_dw"]).max() < 1e-6, tag

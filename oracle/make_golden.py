"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    python oracle/make_golden.py            # rewrites tests/golden/

Every fixture stores the seeded inputs together with what the unmodified reference code
returned, so tests can replay them against (a) the oracle restatement and (b) the CUDA path.
Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch import nn

REF = os.environ.get("SM3_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from src.models import resnet as ref_resnet          # noqa: E402  (reference)
from src.models import simclr as ref_simclr          # noqa: E402  (reference)
from src.models.evaluator import KNNOnlineEvaluator  # noqa: E402  (reference)
from tiny_encoder import tiny                        # noqa: E402

SEED = 3407  # the reference's default seed (src/utils/misc.py:193)


def _np(t):
    return t.detach().cpu().numpy()


def ref_infonce(p1, p2, T, dtype):
    """Unbound SimCLRSkinV3._cal_logits (self unused, simclr.py:290-322) + stock CE + backward."""
    a = p1.to(dtype).clone().requires_grad_(True)
    b = p2.to(dtype).clone().requires_grad_(True)
    logits, labels = ref_simclr.SimCLRSkinV3._cal_logits(None, a, b, nn.Identity(), nn.Identity(), T)
    loss = nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    return logits, labels, loss, a.grad, b.grad


def gen_infonce(name, n, d, T, correlated, store_logits):
    g = torch.Generator().manual_seed(SEED + n + d)
    p1 = torch.randn(n, d, generator=g)
    p2 = p1 + 0.5 * torch.randn(n, d, generator=g) if correlated else torch.randn(n, d, generator=g)
    out = {"p1": _np(p1), "p2": _np(p2), "temperature": np.float64(T), "n": n, "d": d}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        logits, labels, loss, g1, g2 = ref_infonce(p1, p2, T, dt)
        out[f"loss_{tag}"] = _np(loss)
        if store_logits:
            out[f"dp1_{tag}"] = _np(g1)
            out[f"dp2_{tag}"] = _np(g2)
        elif tag == "f64":      # big case: keep a strided subset of gradient rows (fixture size)
            out["grad_rows"] = np.arange(0, n, 5)
            out["dp1_rows_f64"] = _np(g1)[::5]
            out["dp2_rows_f64"] = _np(g2)[::5]
            out["dp1_sum_f64"] = _np(g1).sum(0); out["dp2_sum_f64"] = _np(g2).sum(0)
            out["dp_abs_sum_f64"] = np.float64(_np(g1.abs().sum() + g2.abs().sum()))
        if store_logits:
            out[f"logits_{tag}"] = _np(logits)
        else:
            out[f"logits_row0_{tag}"] = _np(logits[0])
            out[f"logits_col0_{tag}"] = _np(logits[:, 0])
        assert int(labels.abs().sum()) == 0 and labels.dtype == torch.long
        # retrieval contract: top-1 / top-5 non-self neighbour of every row (reference sim matrix)
        if tag == "f64":
            z = torch.nn.functional.normalize(torch.cat([p1, p2]).to(dt), dim=1)
            sim = z @ z.T
            sim.fill_diagonal_(-float("inf"))
            out["top5_idx"] = _np(sim.topk(5, dim=1).indices)
            out["top5_val"] = _np(sim.topk(5, dim=1).values)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", out["loss_f64"])


def gen_edge_cases():
    """Zero row (eps clamp path), duplicated rows, n=1 (no negatives), n=2."""
    out = {}
    g = torch.Generator().manual_seed(SEED)
    cases = {}
    p1 = torch.randn(5, 8, generator=g); p2 = torch.randn(5, 8, generator=g)
    p1[2] = 0.0                                   # F.normalize eps clamp -> z = 0
    cases["zero_row"] = (p1, p2)
    p1 = torch.randn(4, 8, generator=g); p2 = p1.clone()      # z_{i+N} == z_i
    cases["dup_rows"] = (p1, p2)
    cases["n1"] = (torch.randn(1, 8, generator=g), torch.randn(1, 8, generator=g))
    cases["n2"] = (torch.randn(2, 8, generator=g), torch.randn(2, 8, generator=g))
    cases["ragged_n3_d5"] = (torch.randn(3, 5, generator=g), torch.randn(3, 5, generator=g))
    for k, (a, b) in cases.items():
        for T in (0.1, 0.5):
            logits, labels, loss, g1, g2 = ref_infonce(a, b, T, torch.float64)
            key = f"{k}_T{T}"
            out[key + "_p1"] = _np(a); out[key + "_p2"] = _np(b)
            out[key + "_logits"] = _np(logits); out[key + "_loss"] = _np(loss)
            out[key + "_dp1"] = _np(g1); out[key + "_dp2"] = _np(g2)
    out["case_names"] = np.array(sorted(cases.keys()))
    np.savez_compressed(os.path.join(OUT, "infonce_edge.npz"), **out)
    print("edge cases", list(cases))


def gen_heads():
    """8-head CE exactly as the reference scripts write it (they are loops in scripts, not
    functions, so the loops are replayed literally with the reference's criterion objects)."""
    num_classes = [5, 3, 2, 3, 3, 3, 3, 2]           # tools/mlc_eval.py:63
    g = torch.Generator().manual_seed(SEED)
    B = 37
    out = {"num_classes": np.array(num_classes)}
    logits = [torch.randn(B, c, generator=g, dtype=torch.float64).requires_grad_(True) for c in num_classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in num_classes], dim=1)
    # --- tools/mlc_eval.py:157-162: weighted sum / num_labels ---
    label_weights = [1.0, 0.5, 2.0, 1.0, 1.0, 0.25, 1.0, 3.0]
    criterion = nn.CrossEntropyLoss()                # tools/mlc_eval.py:420 region (plain CE)
    loss = 0.0
    for i in range(8):
        loss += label_weights[i] * criterion(logits[i], labels[:, i])
    loss = loss / 8
    loss.backward()
    out["eval_logits"] = np.concatenate([_np(x) for x in logits], axis=1)
    out["eval_labels"] = _np(labels)
    out["eval_weights"] = np.array(label_weights)
    out["eval_loss"] = _np(loss)
    out["eval_grad"] = np.concatenate([_np(x.grad) for x in logits], axis=1)
    # --- tools/mlc_train.py:255-261,381: pred / T, ignore_index=-100, / len(assignments) ---
    T = 0.1
    preds = [torch.randn(B, c, generator=g, dtype=torch.float64).requires_grad_(True) for c in num_classes]
    assign = labels.clone()
    assign[::5, 3] = -100
    assign[1::7, 0] = -100
    criterion = nn.CrossEntropyLoss(ignore_index=-100)   # tools/mlc_train.py:381
    loss = 0
    for i, pred in enumerate(preds):
        scores = pred / T
        loss += criterion(scores, assign[:, i])
    loss = loss / 8
    loss.backward()
    out["dc_logits"] = np.concatenate([_np(x) for x in preds], axis=1)
    out["dc_labels"] = _np(assign)
    out["dc_temperature"] = np.float64(T)
    out["dc_loss"] = _np(loss)
    out["dc_grad"] = np.concatenate([_np(x.grad) for x in preds], axis=1)
    np.savez_compressed(os.path.join(OUT, "heads.npz"), **out)
    print("heads eval loss", out["eval_loss"], "dc loss", out["dc_loss"])


def gen_knn():
    g = torch.Generator().manual_seed(SEED)
    B, NB, D, C, K = 9, 300, 32, 5, 20
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g, dtype=torch.float64), dim=1)
    bank = torch.nn.functional.normalize(torch.randn(NB, D, generator=g, dtype=torch.float64), dim=1)
    tb = torch.randint(0, C, (NB,), generator=g)
    ev = KNNOnlineEvaluator(None, None, C, k=K, temperature=0.07)
    pred = ev.predict(q, bank, tb)                    # evaluator.py:43-83
    sim = q @ bank.T
    w, idx = sim.topk(k=K, dim=-1)                    # evaluator.py:61-63
    np.savez_compressed(os.path.join(OUT, "knn.npz"), query=_np(q), bank=_np(bank), bank_labels=_np(tb),
                        k=K, num_classes=C, temperature=0.07, pred_labels=_np(pred),
                        topk_idx=_np(idx), topk_val=_np(w))
    print("knn pred[:,0]", _np(pred[:, 0]))


def gen_model():
    """Full reference wrappers (SimCLR / SimCLRSkinV3 / SimCLRSkinV32) with the tiny encoder."""
    ref_resnet.__dict__["tiny"] = tiny
    out = {}
    N, PD, T = 6, 8, 0.1
    g = torch.Generator().manual_seed(SEED)
    imgs = [torch.randn(N, 3, 16, 16, generator=g, dtype=torch.float64) for _ in range(4)]
    out["imgs"] = np.stack([_np(x) for x in imgs])
    for cls_name in ("SimCLRSkinV3", "SimCLRSkinV32"):
        torch.manual_seed(SEED)
        model = getattr(ref_simclr, cls_name)("tiny", weights=None, proj_dim=PD, temperature=T).double()
        model.train()
        sd = {k: _np(v) for k, v in model.state_dict().items()}
        for k, v in sd.items():
            out[f"{cls_name}/sd/{k}"] = v
        out[f"{cls_name}/sd_keys"] = np.array(list(sd.keys()))
        criterion = nn.CrossEntropyLoss()
        for style, wts in ((0, (0.5, 0.5)), (1, (0.5, 0.5)), (2, (0.25,) * 4)):
            model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
            model.zero_grad(set_to_none=True)
            outputs = model([imgs[0], imgs[1]], [imgs[2], imgs[3]], style)
            # loss combination of tools/backbone_train.py:98-121
            cross = sum(w * criterion(*outputs[2][k]) for k, w in enumerate(wts))
            derm = criterion(*outputs[0]); clinic = criterion(*outputs[1])
            loss = derm + clinic + cross
            loss.backward()
            pre = f"{cls_name}/style{style}/"
            out[pre + "derm_loss"] = _np(derm); out[pre + "clinic_loss"] = _np(clinic)
            out[pre + "cross_losses"] = np.array([float(criterion(*o)) for o in outputs[2]])
            out[pre + "loss"] = _np(loss)
            out[pre + "logits_shape"] = np.array(outputs[0][0].shape)
            for k, p in model.named_parameters():
                out[pre + "grad/" + k] = _np(p.grad) if p.grad is not None else np.zeros(0)
    out["N"] = N; out["proj_dim"] = PD; out["temperature"] = T
    np.savez_compressed(os.path.join(OUT, "model_tiny.npz"), **out)
    print("model goldens written")


def gen_kmeans():
    """tools/mlc_train.py::cluster_memory (the REAL function) on CPU: a 1-process gloo group stands in for the job and
    Tensor.cuda is made the identity (the function hard-codes .cuda()).  The random initial centroids come from
    torch.randperm on the default generator (:144), so re-seeding reproduces the indices for the fixture."""
    import importlib
    import types
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
        created = True
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        mt = importlib.import_module("tools.mlc_train")
        args = types.SimpleNamespace(world_size=1, rank=0)
        out = {}
        # e / f: the shapes the fused k-means kernel takes (D % 128 == 0, K <= 8): Derm7pt bank size x mlc_proj_dim 512
        cases = [("a", 413, 64, 5, 0.35), ("b", 640, 256, 3, 0.6), ("c", 96, 16, 2, 0.5), ("d", 24, 8, 6, 0.05),
                 ("e", 413, 512, 5, 0.8), ("f", 1000, 128, 8, 0.7)]
        for tag, n, d, k, noise in cases:
            g = torch.Generator().manual_seed(SEED + n + d + k)
            true_c = nn.functional.normalize(torch.randn(k if tag != "d" else 3, d, generator=g), dim=1)
            lab = torch.randint(0, true_c.shape[0], (n,), generator=g)
            emb = nn.functional.normalize(true_c[lab] + noise * torch.randn(n, d, generator=g), dim=1)
            if tag == "d":                      # duplicates of three points -> clusters run empty (the `mask` path :173)
                emb = nn.functional.normalize(true_c[lab], dim=1)
            index = torch.randperm(n, generator=g)
            proto = nn.Linear(d, k, bias=False)
            torch.manual_seed(SEED + 17 * k)
            with torch.no_grad():
                assign = mt.cluster_memory(args, proto, k, index, emb)
            torch.manual_seed(SEED + 17 * k)
            init_idx = torch.randperm(n)[:k]
            out[f"{tag}_emb"] = _np(emb)
            out[f"{tag}_index"] = _np(index)
            out[f"{tag}_init_idx"] = _np(init_idx)
            out[f"{tag}_assign"] = _np(assign)
            out[f"{tag}_centroids"] = _np(proto.weight)
            out[f"{tag}_seed"] = np.int64(SEED + 17 * k)
        out["cases"] = np.array([c[0] for c in cases])
        np.savez_compressed(os.path.join(OUT, "kmeans.npz"), **out)
        print("kmeans.npz", {c[0]: np.bincount(out[f"{c[0]}_assign"][out[f"{c[0]}_assign"] >= 0]).tolist() for c in cases})
    finally:
        torch.Tensor.cuda = orig_cuda
        if created:
            dist.destroy_process_group()


def gen_projtail():
    """The REAL make_projector (src/models/simclr.py:17-27): its last two layers in train mode, then F.normalize as in
    _cal_logits (:294), a random linear functional as the loss; values and autograd gradients in fp64, inputs rounded to
    bf16 so the CUDA path sees identical numbers.  Also the complete term: _cal_logits with two real projectors + CE."""
    from src.models import simclr
    torch.manual_seed(SEED)              # make_projector draws its weights from the default generator
    out = {}
    cases = [("a", 96, 128, 64), ("b", 136, 192, 128), ("c", 160, 64, 256)]
    for tag, r, k, d in cases:
        g = torch.Generator().manual_seed(SEED + r + k + d)
        proj = simclr.make_projector(k, d).double()
        lin, bn = proj[6], proj[7]
        with torch.no_grad():
            lin.weight.copy_(lin.weight.bfloat16().double())
            bn.running_mean.copy_(torch.randn(d, generator=g).double() * 0.1)
            bn.running_var.copy_(1 + 0.2 * torch.rand(d, generator=g).double())
        h = torch.relu(torch.randn(r, k, generator=g)).bfloat16().double().requires_grad_(True)
        gz = torch.randn(r, d, generator=g).double()
        out[f"{tag}_rm0"], out[f"{tag}_rv0"] = _np(bn.running_mean).copy(), _np(bn.running_var).copy()   # before the update
        proj.train()
        z = nn.functional.normalize(bn(lin(h)), dim=1)
        (z * gz).sum().backward()
        out.update({f"{tag}_h": _np(h).astype(np.float32), f"{tag}_w": _np(lin.weight).astype(np.float32),   # bf16 values: exact
                    f"{tag}_gz": _np(gz), f"{tag}_z": _np(z),
                    f"{tag}_dh": _np(h.grad), f"{tag}_dw": _np(lin.weight.grad), f"{tag}_rm1": _np(bn.running_mean),
                    f"{tag}_rv1": _np(bn.running_var)})
        proj.eval()
        out[f"{tag}_z_eval"] = _np(nn.functional.normalize(bn(lin(h.detach())), dim=1))
    # the whole cross-modal term through two real projector tails (V32: one projector per modality, :405-410)
    n, k, d, T = 128, 64, 64, 0.1
    g = torch.Generator().manual_seed(SEED + 777)
    p1, p2 = simclr.make_projector(k, d).double().train(), simclr.make_projector(k, d).double().train()
    tails = [nn.Sequential(p[6], p[7]) for p in (p1, p2)]
    with torch.no_grad():
        for t in tails:
            t[0].weight.copy_(t[0].weight.bfloat16().double())
    f1 = torch.relu(torch.randn(n, k, generator=g)).bfloat16().double().requires_grad_(True)
    f2 = (f1.detach() + 0.3 * torch.relu(torch.randn(n, k, generator=g))).bfloat16().double().requires_grad_(True)
    logits, labels = simclr.SimCLRSkinV3._cal_logits(None, f1, f2, tails[0], tails[1], T)
    loss = nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    out.update({"term_f1": _np(f1).astype(np.float32), "term_f2": _np(f2).astype(np.float32),
                "term_w1": _np(tails[0][0].weight).astype(np.float32), "term_w2": _np(tails[1][0].weight).astype(np.float32),
                "term_loss": _np(loss), "term_df1": _np(f1.grad), "term_df2": _np(f2.grad),
                "term_dw1": _np(tails[0][0].weight.grad), "term_dw2": _np(tails[1][0].weight.grad),
                "term_T": np.float64(T), "term_n": np.int64(n)})
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "projtail.npz"), **out)
    print("projtail.npz", float(loss))


def gen_mlchead():
    """The REAL multi-label ``Model`` (tools/mlc_train.py:58-89) with a stub extractor, MultiLabelProjector4 (run.sh's
    ``--mlc-proj v4``) or Identity projectors, its own TransformerEncoderLayer and prototypes, followed by the DeepCluster
    loss loop of :255-261 (``CE(pred / T, pseudo-label, ignore_index=-100)``, mean over the 8 heads).  Recorded: the
    self-attention output (input of the fused tail, captured by a forward hook), the returned features and predictions,
    the loss, and the gradients that reach the self-attention output and the prototype weights."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_mlc_train", os.path.join(REF, "tools", "mlc_train.py"))
    mt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mt)
    from src.models.projector import MultiLabelProjector4

    class StubExtractor(nn.Module):
        def extract(self, a, b):
            return [a, b]

    out = {}
    cases = [("v4", 24, 48, 128, True, 8), ("v4raw", 20, 32, 256, False, 8), ("ident", 16, 64, 128, True, 1)]
    for tag, b, half, d, l2, hf in cases:
        torch.manual_seed(SEED + b + d)
        feat_dim = 2 * half
        proj = MultiLabelProjector4(feat_dim, d, 8) if hf == 8 else nn.Identity()
        model = mt.Model(StubExtractor(), proj, d if hf == 8 else feat_dim, l2, 1, 64, 0.0).double()
        model.train()
        g = torch.Generator().manual_seed(SEED + 7 * b)
        xa = torch.randn(b, half, generator=g, dtype=torch.float64)
        xb = torch.randn(b, half, generator=g, dtype=torch.float64)
        grabbed = {}

        def grab(_m, _inp, res):
            if res.requires_grad:
                res.retain_grad()
            grabbed["sa"] = res
            grabbed["sa_value"] = res.detach().clone()      # the reference normalises this tensor in place afterwards

        h = model.mlc_sa.register_forward_hook(grab)
        T = 0.7
        targets = torch.stack([torch.randint(0, n, (b,), generator=g) for n in mt.NUM_CLASSES], dim=1)
        targets[::5, 2] = -100
        crit = nn.CrossEntropyLoss(ignore_index=-100)

        def deepcluster_loss(preds):
            loss = 0
            for pred, tgt in zip(preds, targets.t()):
                loss = loss + crit(pred / T, tgt)
            return loss / len(preds)

        if not l2:
            sa_out, preds = model(xa, xb)
            loss = deepcluster_loss(preds)
            loss.backward()
            d_sa, dw = grabbed["sa"].grad, torch.cat([p.weight.grad for p in model.prototypes], dim=0)
        else:
            # With l2_norm the reference normalises `sa_feats[i]` IN PLACE (tools/mlc_train.py:81-83), which autograd
            # rejects in backward ("modified by an inplace operation"), so that path only ever runs under no_grad
            # (init_memory, :96-108).  Forward values from the real Model; gradients from the same arithmetic written
            # out of place on the captured self-attention output.
            with torch.no_grad():
                sa_out, preds = model(xa, xb)
            loss = deepcluster_loss(preds)
            sa_in = grabbed["sa_value"].clone().requires_grad_(True)
            zn = nn.functional.normalize(sa_in, dim=-1, p=2)
            preds2 = [model.prototypes[i](zn[i % len(zn)]) for i in range(len(model.prototypes))]
            assert all(torch.equal(a, b_) for a, b_ in zip(preds, preds2))
            loss2 = deepcluster_loss(preds2)
            loss2.backward()
            d_sa, dw = sa_in.grad, torch.cat([p.weight.grad for p in model.prototypes], dim=0)
        h.remove()
        f32 = lambda t: _np(t).astype(np.float32)           # noqa: E731  (fixture size; compared at 1e-5)
        out.update({f"{tag}_sa_in": f32(grabbed["sa_value"]), f"{tag}_sa_out": f32(sa_out),
                    f"{tag}_logits": _np(torch.cat(preds, dim=1)), f"{tag}_loss": _np(loss),
                    f"{tag}_targets": targets.numpy().astype(np.int64), f"{tag}_T": np.float64(T),
                    f"{tag}_l2": np.int64(int(l2)), f"{tag}_d_sa_in": f32(d_sa),
                    f"{tag}_w": _np(torch.cat([p.weight for p in model.prototypes], dim=0)), f"{tag}_dw": _np(dw)})
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "mlchead.npz"), **out)
    print("mlchead.npz", float(loss))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    gen_infonce("infonce_n4_d8_T05", 4, 8, 0.5, False, True)
    gen_infonce("infonce_n64_d128_T01", 64, 128, 0.1, False, True)        # BASELINE config 1
    gen_infonce("infonce_n48_d128_T01_corr", 48, 128, 0.1, True, True)    # run.sh scale: 96/2 GPUs
    gen_infonce("infonce_n200_d64_T02_corr", 200, 64, 0.2, True, False)   # ragged vs 128-row tiles
    gen_infonce("infonce_n512_d256_T01", 512, 256, 0.1, False, False)     # config-4 width
    gen_edge_cases()
    gen_heads()
    gen_knn()
    gen_model()
    gen_kmeans()
    gen_projtail()
    gen_mlchead()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kmeans":      # regenerate one fixture without touching the others
        os.makedirs(OUT, exist_ok=True)
        gen_kmeans()
    elif len(sys.argv) > 1 and sys.argv[1] == "projtail":
        os.makedirs(OUT, exist_ok=True)
        gen_projtail()
    elif len(sys.argv) > 1 and sys.argv[1] == "mlchead":
        os.makedirs(OUT, exist_ok=True)
        gen_mlchead()
    else:
        main()

"""multiclass_auroc / recall / specificity / precision with ``average=None`` (per-class vector) or ``"macro"``."""
import torch


def _labels(preds: torch.Tensor) -> torch.Tensor:
    return preds.argmax(dim=1) if preds.dim() == 2 else preds


def _confusion(preds, target, num_classes):
    p, t = _labels(preds).long(), target.long()
    cm = torch.zeros(num_classes, num_classes, dtype=torch.float64, device=t.device)
    cm.index_put_((t, p), torch.ones_like(t, dtype=torch.float64), accumulate=True)
    tp = cm.diag()
    fn = cm.sum(1) - tp
    fp = cm.sum(0) - tp
    tn = cm.sum() - tp - fn - fp
    return tp, fp, tn, fn


def _reduce(v, average):
    v = v.float()
    if average in (None, "none"):
        return v
    if average == "macro":
        return v.mean()
    raise ValueError(f"average={average!r} is not supported by the sm3 torchmetrics shim")


def _safe_div(a, b):
    return torch.where(b > 0, a / b.clamp(min=1), torch.zeros_like(a))


def multiclass_recall(preds, target, num_classes, average="macro", **_):
    tp, fp, tn, fn = _confusion(preds, target, num_classes)
    return _reduce(_safe_div(tp, tp + fn), average)


def multiclass_specificity(preds, target, num_classes, average="macro", **_):
    tp, fp, tn, fn = _confusion(preds, target, num_classes)
    return _reduce(_safe_div(tn, tn + fp), average)


def multiclass_precision(preds, target, num_classes, average="macro", **_):
    tp, fp, tn, fn = _confusion(preds, target, num_classes)
    return _reduce(_safe_div(tp, tp + fp), average)


def multiclass_auroc(preds, target, num_classes, average="macro", **_):
    """One-vs-rest area under the ROC curve from the rank statistic (ties get the average rank)."""
    scores = preds.double()
    if scores.dim() != 2:
        raise ValueError("multiclass_auroc expects class scores [B, C]")
    if (scores < 0).any() or (scores > 1).any():
        scores = scores.softmax(dim=1)                     # torchmetrics normalises logits the same way
    out = torch.zeros(num_classes, dtype=torch.float64, device=scores.device)
    for c in range(num_classes):
        s = scores[:, c]
        pos = target == c
        n_pos, n_neg = int(pos.sum()), int((~pos).sum())
        if n_pos == 0 or n_neg == 0:
            continue                                        # undefined: torchmetrics reports 0 with a warning
        order = torch.argsort(s)
        ranks = torch.empty_like(s)
        ranks[order] = torch.arange(1, s.numel() + 1, dtype=torch.float64, device=s.device)
        uniq, inv = torch.unique(s, return_inverse=True)    # average the ranks of tied scores
        sums = torch.zeros_like(uniq).index_add_(0, inv, ranks)
        cnts = torch.zeros_like(uniq).index_add_(0, inv, torch.ones_like(ranks))
        ranks = (sums / cnts)[inv]
        out[c] = (ranks[pos].sum() - n_pos * (n_pos + 1) / 2) / (n_pos * n_neg)
    return _reduce(out, average)

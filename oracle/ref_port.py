"""Torch-CPU port of the reference's materialising hot path  --  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python on top of torch and cannot travel to the GPU box
(/root/reference does not exist there), so the CPU baseline that bench.py times beside the
CUDA path is this op-for-op port: it issues the same ATen operator sequence as
src/models/simclr.py:290-322 followed by nn.CrossEntropyLoss (tools/backbone_train.py:531),
so its cost profile (nonzero / index / index_put / mm / log_softmax) is the reference's.
``tests/test_oracle_golden.py`` pins it to outputs of the real reference (tests/golden/).

Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def port_cal_logits(f1: torch.Tensor, f2: torch.Tensor, temperature: float):
    """(logits[M, M-1], labels[M]) exactly as the reference builds them after the projectors.

    Op sequence kept identical to simclr.py:293-320: cat -> normalize -> pair-id equality matrix
    (built on the host, then moved) -> Gram matrix -> drop diagonal by boolean mask ->
    boolean-mask gather of positives / negatives -> cat -> divide by temperature.
    """
    n = f1.shape[0]
    dev = f1.device
    feats = F.normalize(torch.cat((f1, f2), dim=0), dim=1)             # :293-294
    pair_id = torch.arange(n).repeat(2)                                # :296
    same_pair = pair_id[None, :].eq(pair_id[:, None]).float().to(dev)  # :297-298
    gram = feats @ feats.T                                             # :300
    m = same_pair.shape[0]
    keep = ~torch.eye(m, dtype=torch.bool).to(dev)                     # :303
    same_pair = same_pair[keep].view(m, -1)                            # :304
    gram = gram[keep].view(m, -1)                                      # :305-307
    pos = gram[same_pair.bool()].view(m, -1)                           # :310
    neg = gram[~same_pair.bool()].view(m, -1)                          # :313-315
    logits = torch.cat((pos, neg), dim=1) / temperature                # :317,320
    target = torch.zeros(m, dtype=torch.long).to(dev)                  # :318
    return logits, target


def port_infonce_step(p1: torch.Tensor, p2: torch.Tensor, temperature: float):
    """One InfoNCE term forward + backward the way tools/backbone_train.py:101-125 runs it.

    Returns (loss, dp1, dp2)."""
    a = p1.detach().clone().requires_grad_(True)
    b = p2.detach().clone().requires_grad_(True)
    logits, target = port_cal_logits(a, b, temperature)
    loss = F.cross_entropy(logits, target)                             # nn.CrossEntropyLoss()
    loss.backward()
    return loss.detach(), a.grad, b.grad


def port_infonce_step_rowblock(p1: torch.Tensor, p2: torch.Tensor, temperature: float,
                               row_start: int, row_count: int):
    """Bounded SAMPLE of the same workload for shapes whose [M,M] matrix does not fit in host RAM.

    Runs the reference's operator sequence on a block of `row_count` rows of the Gram matrix
    against ALL M columns (mask, gather positives/negatives, cat, /T, CE sum, backward).  The
    per-row work is identical to the full step; bench.py scales the measured time by
    M / row_count and says so in ``cpu_baseline.sample``.
    """
    n = p1.shape[0]
    m = 2 * n
    a = p1.detach().clone().requires_grad_(True)
    b = p2.detach().clone().requires_grad_(True)
    feats = F.normalize(torch.cat((a, b), dim=0), dim=1)
    rows = torch.arange(row_start, row_start + row_count)
    pair_id = torch.arange(n).repeat(2)
    same_pair = pair_id[None, :].eq(pair_id[rows][:, None]).float()
    gram = feats[rows] @ feats.T
    keep = torch.ones(row_count, m, dtype=torch.bool)
    keep[torch.arange(row_count), rows] = False
    same_pair = same_pair[keep].view(row_count, -1)
    gram = gram[keep].view(row_count, -1)
    pos = gram[same_pair.bool()].view(row_count, -1)
    neg = gram[~same_pair.bool()].view(row_count, -1)
    logits = torch.cat((pos, neg), dim=1) / temperature
    target = torch.zeros(row_count, dtype=torch.long)
    loss = F.cross_entropy(logits, target, reduction="sum") / m
    loss.backward()
    return loss.detach(), a.grad, b.grad


def port_multihead_ce(outputs, labels: torch.Tensor, weights=None, temperature: float = 1.0,
                      ignore_index: int = -100):
    """The 8-head loop of tools/mlc_eval.py:159-162 (weights) / tools/mlc_train.py:255-261 (pred/T)."""
    h = len(outputs)
    total = 0.0
    for i in range(h):
        w = 1.0 if weights is None else float(weights[i])
        total = total + w * F.cross_entropy(outputs[i] / temperature, labels[:, i], ignore_index=ignore_index)
    return total / h

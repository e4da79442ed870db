"""Drop-in ``src.models.simclr`` on the fused sm_100a kernels.

Same public names, constructor signatures, attribute names, ``state_dict`` keys and return structure as the
reference module (src/models/simclr.py), so tools/backbone_train.py:39,489-507 and tools/mlc_train.py:33,337-346
run unchanged.  What differs is what happens between the projector and the loss: instead of building the
``[2N, 2N-1]`` logits matrix (reference :62-88, :138-164, :294-320) each term returns the sufficient-statistics
logits ``[2N, 2]`` from ``skin_sm3_b200.cal_logits``; the script's own ``nn.CrossEntropyLoss()`` on them is the
reference loss, value and gradient.

Runtime switches (environment, read at call time so unchanged scripts can opt in):
  SM3_PRECISION         auto | bf16 | fp32     (default auto: fp32 inputs outside autocast stay exact fp32)
  SM3_GLOBAL_NEGATIVES  1 -> all-gather the normalised embeddings over the default process group
  SM3_FUSED_TAIL        0 -> keep the projector's last Linear + BatchNorm on stock kernels (default 1: under autocast the
                        last two layers of make_projector and the F.normalize after them run as the fused tail,
                        skin_sm3_b200.tail_cal_logits; fp32 inputs always take the stock layers)
"""
import os
from copy import deepcopy
from functools import partial

import torch
import torch.nn as nn

from skin_sm3_b200 import functional as F3
from src.models import resnet


def make_projector(in_dim, proj_dim):
    # layer order / indices fix the state_dict keys (0,1,3,4,6,7) -- reference simclr.py:17-27
    layers = []
    for _ in range(2):
        layers += [nn.Linear(in_dim, in_dim, bias=False), nn.BatchNorm1d(in_dim), nn.ReLU(inplace=True)]
    layers += [nn.Linear(in_dim, proj_dim, bias=False), nn.BatchNorm1d(proj_dim, affine=False)]
    return nn.Sequential(*layers)


def _runtime():
    group = None
    if os.environ.get("SM3_GLOBAL_NEGATIVES", "0") == "1" and torch.distributed.is_available() \
            and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        group = torch.distributed.group.WORLD
    return os.environ.get("SM3_PRECISION", "auto"), group


def _check_range(proj_dim, temperature):
    """The fused kernels take D <= 256 and 1/T < 83 (include/sm3_b200.h); say so when the model is BUILT, not at the first
    forward.  (The reference accepts any --proj-dim / --temperature; its configurations use 128 / 0.1.)"""
    if int(proj_dim) > 256 or int(proj_dim) < 1:
        raise ValueError(f"skin_sm3_b200 drop-in: proj_dim={proj_dim} is outside the fused kernels' range (1..256); "
                         f"run with SM3_DROPIN=0 for the stock reference modules")
    if not (float(temperature) > 0.0 and 1.0 / float(temperature) < 83.0):
        raise ValueError(f"skin_sm3_b200 drop-in: temperature={temperature} is outside the fused kernels' range "
                         f"(T > 0.012); run with SM3_DROPIN=0 for the stock reference modules")


def _pair_logits(feats_a, feats_b, temperature):
    """(logits[2N,2], zeros[2N]) for the pairing rows(feats_a)[i] <-> rows(feats_b)[i]."""
    precision, group = _runtime()
    return F3.cal_logits(feats_a, feats_b, temperature, precision, group)


def _projected_logits(groups, temperature):
    """groups: [(x, projector)] -- one entry holding all 2N rows (both views through one projector, reference :61) or two
    entries of N rows (one projector per half, reference :293).  Under autocast the projector's last Linear + affine-free
    BatchNorm + the F.normalize that follows go through the fused tail; otherwise the stock layers + cal_logits."""
    precision, group = _runtime()
    fused = os.environ.get("SM3_FUSED_TAIL", "1") != "0" and precision != "fp32"
    tails = []
    if fused:
        for x, proj in groups:
            if not (isinstance(proj, nn.Sequential) and len(proj) == 8):
                fused = False
                break
            h = proj[:6](x)
            if not F3.tail_supported(h, proj[6], proj[7]):
                tails.append((None, proj[6:](h)))          # the body has run; finish this group on the stock layers
                fused = False
            else:
                tails.append((F3.TailSpec(h, proj[6].weight.to(h.dtype), proj[7]), None))
        if fused:
            return F3.tail_cal_logits([t for t, _ in tails], temperature, group)
        # mixed / unsupported: fall through with what has been computed so far
        outs = [p if p is not None else t.bn(torch.nn.functional.linear(t.h, t.w)) for t, p in tails]
        outs += [proj(x) for x, proj in groups[len(tails):]]
    else:
        outs = [proj(x) for x, proj in groups]
    if len(outs) == 1:
        n = outs[0].shape[0] // 2
        return F3.cal_logits(outs[0][:n], outs[0][n:], temperature, precision, group)
    return F3.cal_logits(outs[0], outs[1], temperature, precision, group)


class SimCLR(nn.Module):
    """One modality: encoder + projector, intra-modal InfoNCE between two augmented views (reference :31-96)."""

    def __init__(self, arch, weights=None, proj_dim=128, temperature=0.5, return_feats=False):
        super().__init__()
        _check_range(proj_dim, temperature)
        self.proj_dim = proj_dim
        self.temperature = temperature
        self.return_feats = return_feats
        self.encoder = resnet.__dict__[arch](weights=weights)
        self.encoder_out_dim = self.encoder.fc.in_features
        self.encoder.fc = nn.Identity()
        self.projector = make_projector(self.encoder_out_dim, self.proj_dim)

    def forward(self, x1, x2):
        n = x1.shape[0]
        f1 = self.encoder(x1)
        f2 = self.encoder(x2)
        # one projector pass over both views: its BatchNorm sees all 2N rows jointly (reference :61)
        out = _projected_logits([(torch.cat([f1, f2], dim=0), self.projector)], self.temperature)
        return (out, (f1, f2)) if self.return_feats else out

    def extract(self, imgs):
        return self.encoder(imgs)


class _TwoBranch(nn.Module):
    def extract(self, derm_imgs, clinic_imgs):
        return [self.derm_backbone.encoder(derm_imgs), self.clinic_backbone.encoder(clinic_imgs)]


class SimCLRSkin(_TwoBranch):
    def __init__(self, arch, weights=None, proj_dim=128, temperature=0.5) -> None:
        super().__init__()
        self.derm_backbone = SimCLR(arch, weights, proj_dim, temperature)
        self.clinic_backbone = SimCLR(arch, weights, proj_dim, temperature)

    def forward(self, derm_imgs, clinic_imgs):
        return (self.derm_backbone(*derm_imgs), self.clinic_backbone(*clinic_imgs))


class SimCLRSkinV2(_TwoBranch):
    """Fusion-by-concat ablations: one projector over cat([f1, f2]) (reference :118-183)."""

    def __init__(self, arch, weights=None, proj_dim=128, temperature=0.5) -> None:
        super().__init__()
        self.temperature = temperature
        self.derm_backbone = SimCLR(arch, weights, proj_dim, temperature, True)
        self.clinic_backbone = SimCLR(arch, weights, proj_dim, temperature, True)
        cross_feat_dim = self.derm_backbone.encoder_out_dim + self.clinic_backbone.encoder_out_dim
        self.cross_proj = make_projector(cross_feat_dim, proj_dim)

    def _cal_logits(self, f1, f2, projector, temperature):
        n = f1.shape[0]
        proj = projector(torch.cat([f1, f2], dim=0))
        return _pair_logits(proj[:n], proj[n:], temperature)

    def _branches(self, derm_imgs, clinic_imgs):
        derm_outs, derm_feats = self.derm_backbone(*derm_imgs)
        clinic_outs, clinic_feats = self.clinic_backbone(*clinic_imgs)
        return derm_outs, clinic_outs, derm_feats, clinic_feats

    def forward(self, derm_imgs, clinic_imgs):
        derm_outs, clinic_outs, d, c = self._branches(derm_imgs, clinic_imgs)
        cross = self._cal_logits(torch.cat([d[0], c[0]], dim=1), torch.cat([d[1], c[1]], dim=1),
                                 self.cross_proj, self.temperature)
        return (derm_outs, clinic_outs, cross)


class SimCLRSkinV21(SimCLRSkinV2):
    def forward(self, derm_imgs, clinic_imgs):
        derm_outs, clinic_outs, d, c = self._branches(derm_imgs, clinic_imgs)
        cross = self._cal_logits(torch.cat([d[0], c[1]], dim=1), torch.cat([d[1], c[0]], dim=1),
                                 self.cross_proj, self.temperature)
        return (derm_outs, clinic_outs, cross)


class SimCLRSkinV22(SimCLRSkinV2):
    def forward(self, derm_imgs, clinic_imgs):
        derm_outs, clinic_outs, d, c = self._branches(derm_imgs, clinic_imgs)
        straight = self._cal_logits(torch.cat([d[0], c[0]], dim=1), torch.cat([d[1], c[1]], dim=1),
                                    self.cross_proj, self.temperature)
        crossed = self._cal_logits(torch.cat([d[0], c[1]], dim=1), torch.cat([d[1], c[0]], dim=1),
                                   self.cross_proj, self.temperature)
        return (derm_outs, clinic_outs, (straight, crossed))


class SimCLRSkinV23(SimCLRSkinV2):
    def __init__(self, arch, weights=None, proj_dim=128, temperature=0.5) -> None:
        super().__init__(arch, weights, proj_dim, temperature)
        self.cross_proj = make_projector(self.derm_backbone.encoder_out_dim, proj_dim)

    def forward(self, derm_imgs, clinic_imgs):
        derm_outs, clinic_outs, d, c = self._branches(derm_imgs, clinic_imgs)
        cross = self._cal_logits(d[0] + c[0], d[1] + c[1], self.cross_proj, self.temperature)
        return (derm_outs, clinic_outs, cross)


# which (derm view, clinic view) pairs are contrasted per `style` (reference :328-389 / :419-480)
_STYLE_PAIRS = {0: ((0, 0), (1, 1)), 1: ((0, 1), (1, 0)), 2: ((0, 0), (0, 1), (1, 0), (1, 1))}


class SimCLRSkinV3(_TwoBranch):
    """Cross-modal contrast derm <-> clinic with one shared projector (reference :250-396)."""

    def __init__(self, arch, weights=None, proj_dim=128, temperature=0.5, use_checkpoint=False) -> None:
        super().__init__()
        self.temperature = temperature
        self.derm_backbone = SimCLR(arch, weights, proj_dim, temperature, True)
        self.clinic_backbone = SimCLR(arch, weights, proj_dim, temperature, True)
        self.derm_feat_dim = self.derm_backbone.encoder_out_dim
        self.clinic_feat_dim = self.clinic_backbone.encoder_out_dim
        self.cross_feat_dim = self.derm_feat_dim
        self.cross_proj = make_projector(self.cross_feat_dim, proj_dim)
        if use_checkpoint:
            self._apply_checkpoint()

    def _apply_checkpoint(self):
        from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import (
            CheckpointImpl, apply_activation_checkpointing, checkpoint_wrapper)
        keep = [deepcopy(self.derm_backbone.encoder.conv1), deepcopy(self.clinic_backbone.encoder.conv1)]
        wrap = partial(checkpoint_wrapper, offload_to_cpu=False, checkpoint_impl=CheckpointImpl.NO_REENTRANT)
        apply_activation_checkpointing(self, checkpoint_wrapper_fn=wrap,
                                       check_fn=lambda m: isinstance(m, (nn.Conv2d, nn.Linear)))
        self.derm_backbone.encoder.conv1, self.clinic_backbone.encoder.conv1 = keep   # first convs stay plain

    def _cal_logits(self, f1, f2, projector1, projector2, temperature):
        # each modality goes through its projector separately: BatchNorm statistics per N rows (reference :293)
        return _projected_logits([(f1, projector1), (f2, projector2)], temperature)

    def _projectors(self):
        return self.cross_proj, self.cross_proj

    def forward(self, derm_imgs, clinic_imgs, style):
        derm_outs, derm_feats = self.derm_backbone(*derm_imgs)
        clinic_outs, clinic_feats = self.clinic_backbone(*clinic_imgs)
        proj_d, proj_c = self._projectors()
        cross_outs = tuple(self._cal_logits(derm_feats[a], clinic_feats[b], proj_d, proj_c, self.temperature)
                           for a, b in _STYLE_PAIRS[style])
        return (derm_outs, clinic_outs, cross_outs)


class SimCLRSkinV32(SimCLRSkinV3):
    """As V3 with an independent projector per modality: ``cross_proj`` is a ModuleList of two (reference :399-482)."""

    def __init__(self, arch, weights=None, proj_dim=128, temperature=0.5, use_checkpoint=False) -> None:
        super().__init__(arch, weights, proj_dim, temperature)
        self.cross_proj = nn.ModuleList([make_projector(self.derm_feat_dim, proj_dim),
                                         make_projector(self.clinic_feat_dim, proj_dim)])
        if use_checkpoint:
            self._apply_checkpoint()

    def _projectors(self):
        return self.cross_proj[0], self.cross_proj[1]

"""Drop-in layer: the reference's module path ``src.models.simclr`` re-implemented on the fused kernels.

``install()`` registers an import hook so that ``from src.models.simclr import SimCLRSkinV3, SimCLRSkinV32``
(tools/backbone_train.py:39, tools/mlc_train.py:33) resolves to the classes defined here; putting
``site_dir()`` on PYTHONPATH does the same for unchanged scripts and their mp.spawn workers.  ``patch_reference()`` is the alternative for a
process that already imported the reference's module: it swaps the two logits builders in place.
"""
import os
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))


def install() -> str:
    """Serve ``src.models.simclr`` from this package in the current interpreter (import hook; independent of
    sys.path order).  For unchanged scripts / spawned workers use PYTHONPATH=<this dir>/_site instead."""
    from ._hook import install_hook
    install_hook()
    return _DIR


def site_dir() -> str:
    """Directory to put on PYTHONPATH (contains the sitecustomize that installs the hook everywhere)."""
    return os.path.join(_DIR, "_site")


def patch_reference(ref_simclr_module, precision: str = "auto", group=None) -> None:
    """Monkey-patch an already-imported reference ``src.models.simclr`` so that SimCLR.forward and every
    ``_cal_logits`` (V2 family :134-166, V3 family :290-322) go through the fused kernels."""
    import torch
    from .. import functional as F3

    def v3_cal_logits(self, f1, f2, projector1, projector2, temperature):
        return F3.cal_logits(projector1(f1), projector2(f2), temperature, precision, group)

    def v2_cal_logits(self, f1, f2, projector, temperature):
        n = f1.shape[0]
        feats = projector(torch.cat([f1, f2], dim=0))
        return F3.cal_logits(feats[:n], feats[n:], temperature, precision, group)

    def simclr_forward(self, x1, x2):
        n = x1.shape[0]
        f1, f2 = self.encoder(x1), self.encoder(x2)
        feats = self.projector(torch.cat([f1, f2], dim=0))
        out = F3.cal_logits(feats[:n], feats[n:], self.temperature, precision, group)
        return (out, (f1, f2)) if self.return_feats else out

    ref_simclr_module.SimCLRSkinV3._cal_logits = v3_cal_logits
    ref_simclr_module.SimCLRSkinV2._cal_logits = v2_cal_logits
    ref_simclr_module.SimCLR.forward = simclr_forward

"""Minimal stand-in for the ``torchmetrics`` names the reference's scripts import at module level
(tools/backbone_train.py:32-37, tools/backbone_eval.py:30, tools/mlc_eval.py:31).  Only served when the real package is
not installed (skin_sm3_b200/dropin/_hook.py checks first).  The four metrics are implemented in plain torch with the
argument meaning of torchmetrics' functional API as the reference calls it (src/utils/misc.py:299-327):
``fn(preds[B, C], target[B], num_classes=C, average=None) -> [C]``."""
__version__ = "0+sm3shim"

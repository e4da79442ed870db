"""Encoder registry of the drop-in ``src.models`` package.

The reference ships a verbatim copy of torchvision 0.13's ResNet (src/models/resnet.py:1-8) and looks
encoders up with ``resnet.__dict__[arch](weights=weights)`` (src/models/simclr.py:47).  The backbone is out
of scope for this project (north_star: "stays on stock PyTorch/cuDNN"), so the same names are re-exported
from the installed torchvision.
"""
from torchvision.models.resnet import (ResNet, resnet18, resnet34, resnet50, resnet101, resnet152,  # noqa: F401
                                       resnext50_32x4d, resnext101_32x8d, wide_resnet50_2, wide_resnet101_2)

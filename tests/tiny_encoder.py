"""A 600-parameter stand-in encoder shared by oracle/make_golden.py and the drop-in tests.

The reference picks its encoder with ``resnet.__dict__[arch](weights=weights)`` and only needs
``.fc.in_features`` (src/models/simclr.py:47-49).  Registering this class under the name
``"tiny"`` in that dict lets the *unmodified* reference wrappers (SimCLR / SimCLRSkinV3 / V32)
run end to end with a state_dict small enough to commit as a golden fixture.
"""
import torch
from torch import nn

TINY_FEAT_DIM = 16


class TinyEncoder(nn.Module):
    def __init__(self, weights=None):
        super().__init__()
        self.conv1 = nn.Conv2d(3, TINY_FEAT_DIM, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(TINY_FEAT_DIM)
        self.relu = nn.ReLU(inplace=True)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(TINY_FEAT_DIM, 10)

    def forward(self, x):
        x = self.avgpool(self.relu(self.bn1(self.conv1(x))))
        return self.fc(torch.flatten(x, 1))


def tiny(weights=None):
    return TinyEncoder(weights)

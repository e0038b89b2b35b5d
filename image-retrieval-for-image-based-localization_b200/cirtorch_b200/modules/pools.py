"""Pooling layers with the reference's names and constructor signatures
(cirtorch/modules/pools.py:10-54, registry :200-207), running the fused CUDA tail kernel."""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from .. import functional as LF


class MAC(nn.Module):
    """cirtorch/modules/pools.py:10-16."""

    def forward(self, x):
        return LF.mac(x)

    def __repr__(self):
        return self.__class__.__name__ + "()"


class SPoC(nn.Module):
    """cirtorch/modules/pools.py:20-26."""

    def forward(self, x):
        return LF.spoc(x)

    def __repr__(self):
        return self.__class__.__name__ + "()"


class GeM(nn.Module):
    """cirtorch/modules/pools.py:30-38: learnable exponent ``p`` (Parameter of shape [1]), eps clamp."""

    def __init__(self, p=3, eps=1e-6):
        super().__init__()
        self.p = Parameter(torch.ones(1) * p)
        self.eps = eps

    def forward(self, x):
        return LF.gem(x, p=self.p, eps=self.eps)

    def __repr__(self):
        return self.__class__.__name__ + "(p=%.4f, eps=%s)" % (float(self.p.detach().flatten()[0]), self.eps)


class GeMmp(nn.Module):
    """cirtorch/modules/pools.py:43-54: one exponent per channel (``mp`` = number of channels)."""

    def __init__(self, p=3, mp=1, eps=1e-6):
        super().__init__()
        self.mp = mp
        self.p = Parameter(torch.ones(self.mp) * p)
        self.eps = eps

    def forward(self, x):
        return LF.descriptor_tail(x, p=self.p, eps=self.eps, pooling="GeMmp", pool_only=True)[:, :, None, None]

    def __repr__(self):
        return self.__class__.__name__ + "(p=[%d], eps=%s)" % (self.mp, self.eps)


# RMAC / ROIpool (pools.py:57-197) are outside the hot path named by BASELINE.json.
POOLING_LAYERS = {
    "MAC": MAC,
    "SPoC": SPoC,
    "GeM": GeM,
    "GeMmp": GeMmp,
}

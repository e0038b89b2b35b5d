"""cirtorch/modules/normalizations.py:9-33 on the CUDA row-L2N kernel."""
import torch.nn as nn

from .. import functional as LF


class L2N(nn.Module):
    """x / (||x||_2 over dim 1 + eps); eps is ADDED to the norm (normalizations.py:15-16)."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return LF.l2n(x, eps=self.eps)

    def __repr__(self):
        return self.__class__.__name__ + "(eps=%s)" % self.eps


class PowerLaw(nn.Module):
    """Signed square root (normalizations.py:19-27)."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return LF.powerlaw(x, eps=self.eps)

    def __repr__(self):
        return self.__class__.__name__ + "(eps=%s)" % self.eps


NORMALIZATION_LAYERS = {
    "L2N": L2N,
    "PowerLaw": PowerLaw,
}

"""cirtorch/modules/normalizations.py:9-33 on the CUDA row kernels: same class names, constructor (eps=1e-6), registry keys."""
import torch.nn as nn

from .. import functional as LF


class _RowNormalization(nn.Module):
    """A normalisation with one parameter-free epsilon; subclasses name the CUDA op."""

    _op = None

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return type(self)._op(x, eps=self.eps)

    def __repr__(self):
        return "%s(eps=%s)" % (type(self).__name__, self.eps)


class L2N(_RowNormalization):
    """x / (||x||_2 over dim 1 + eps); eps is ADDED to the norm (normalizations.py:15-16)."""

    _op = staticmethod(LF.l2n)


class PowerLaw(_RowNormalization):
    """Signed square root of x + eps (normalizations.py:19-27)."""

    _op = staticmethod(LF.powerlaw)


NORMALIZATION_LAYERS = dict(L2N=L2N, PowerLaw=PowerLaw)

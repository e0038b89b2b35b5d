"""Upstream-compatible import paths (cirtorch.layers.pooling / normalization / functional);
the fork keeps these modules empty and defines the layers under cirtorch/modules."""

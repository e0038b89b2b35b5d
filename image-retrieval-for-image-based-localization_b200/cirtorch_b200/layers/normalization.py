from ..modules.normalizations import L2N, NORMALIZATION_LAYERS  # noqa: F401

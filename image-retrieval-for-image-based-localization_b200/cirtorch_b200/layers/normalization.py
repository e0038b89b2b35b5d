from ..modules.normalizations import L2N, PowerLaw, NORMALIZATION_LAYERS  # noqa: F401

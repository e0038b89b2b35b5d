from ..modules.pools import MAC, SPoC, GeM, GeMmp, POOLING_LAYERS  # noqa: F401

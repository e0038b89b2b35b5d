from ..functional import gem, mac, spoc, l2n, powerlaw, descriptor_tail  # noqa: F401

from ..functional import gem, mac, spoc, l2n, descriptor_tail  # noqa: F401

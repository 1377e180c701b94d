from vstnet_b200.cWCT import cWCT  # noqa: F401

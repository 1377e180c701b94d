from vstnet_b200.RevResNet import RevResNet, residual_block, channel_reduction  # noqa: F401

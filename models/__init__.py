"""Drop-in shim: lets the reference's scripts (`from models.RevResNet import RevResNet`,
`from models.cWCT import cWCT`; image_transfer.py:44,59) run unchanged against vstnet_b200."""

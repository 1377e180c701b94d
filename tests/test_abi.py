"""CPU: the C-ABI library loads and exports every symbol include/vstb200.h declares; host-side
argument checking and the drop-in surface of the Python classes."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import ROOT, build_net


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vstb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vst_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vstnet_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "libvstb200.so does not export %s" % s
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table out of sync with the header"
    assert _lib.load().vst_version() >= 100


def test_plan_metadata_without_gpu():
    from vstnet_b200 import _lib
    lib = _lib.load()
    net = build_net("photo")
    assert lib.vst_revnet_latent_channels(net._h) == 32
    assert lib.vst_revnet_down_scale(net._h) == 4
    assert lib.vst_revnet_param_floats(net._h) == 4089936
    assert lib.vst_revnet_workspace_bytes(net._h, 1, 1080, 1920) >= 3 * 16 * 1080 * 1920 * 4
    art = build_net("art")
    assert lib.vst_revnet_latent_channels(art._h) == 128


def test_bad_config_is_rejected():
    from vstnet_b200 import RevResNet
    with pytest.raises(ValueError):
        RevResNet(nChannels=[16, 48, 256])          # stride-2 stage must quadruple the width
    with pytest.raises(ValueError):
        RevResNet(hidden_dim=8, sp_steps=2)         # 8*16 < 256: negative channel_reduction pad


def test_state_dict_surface_matches_reference():
    net = build_net("photo")
    keys = list(net.state_dict().keys())
    assert keys[0] == "stack.0.conv.1.weight" and keys[1] == "stack.0.conv.1.bias"
    assert "stack.29.conv.7.weight" in keys and "channel_reduction.block_list.1.conv.4.bias" in keys
    assert net.state_dict()["stack.10.conv.1.weight"].shape == (16, 16, 3, 3)     # stride-2 block: in_ch = c/4
    assert net.state_dict()["stack.20.conv.7.weight"].shape == (256, 64, 3, 3)
    assert int(net.down_scale) == 4 and net.pad == 29 and net.in_ch == 16 and net.nBlocks == [10, 10, 10]
    # a reference-style checkpoint round-trips through load_state_dict
    other = build_net("photo", seed=1)
    other.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
    assert all(torch.equal(a, b) for a, b in zip(other.state_dict().values(), net.state_dict().values()))


def test_no_cpu_fallback():
    from vstnet_b200 import cWCT
    net = build_net("photo")
    with pytest.raises(RuntimeError, match="no CPU"):
        net(torch.rand(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU"):
        cWCT().transfer(torch.rand(1, 32, 8, 8), torch.rand(1, 32, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU"):
        cWCT().transfer(torch.rand(1, 32, 8, 8), torch.rand(1, 32, 8, 8),
                        np.zeros((1, 8, 8), np.uint8), np.zeros((1, 8, 8), np.uint8))


def test_models_shim_is_the_drop_in_import_path():
    from models.RevResNet import RevResNet as A
    from models.cWCT import cWCT as B
    import vstnet_b200
    assert A is vstnet_b200.RevResNet and B is vstnet_b200.cWCT

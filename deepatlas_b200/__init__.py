"""deepatlas_b200 -- B200-native (sm_100a) implementation of DeepAtlas's volumetric training hot path.

Public surface (mirrors the reference's two registries):

    from deepatlas_b200 import get_network, get_loss_function, install
    install()                       # overwrite lib.network_factory.network_dic / lib.loss.loss_dict entries
    net = get_network('UNet_light')(1, 32, bias=True, BN=True).cuda()
    crit = get_loss_function('dice')(n_class=32, weight_type='Uniform', softmax=True, eps=1e-6)

Everything numerical runs in ``libdeepatlas_b200.so`` (hand-written CUDA behind a C ABI, see
include/deepatlas_b200.h); importing this package without the built library raises on first use.
"""
from __future__ import annotations

from . import _lib, evaluation, ops  # noqa: F401
from .losses import (BendingEnergyLoss, CrossEntropyLoss, DiceLossMultiClass, DiceLossOnLabel, FocalLoss,  # noqa: F401
                     L2Loss, LNCCLoss, MSELoss, NormalizedCrossCorrelationLoss, SoftCrossEntropy, VoxelMorphLNCC,
                     get_available_losses, get_loss_function, gradientLoss, loss_dict)
from .networks import (UNet, UNet_generator, UNet_light, VoxelMorphCVPR2018, get_available_networks,  # noqa: F401
                       get_network, network_dic)

__version__ = "0.1.0"


def install(network_factory_module=None, loss_module=None):
    """Install the B200 classes into the REFERENCE's own registries (the sanctioned seams):
    ``lib.network_factory.network_dic`` (lib/network_factory/__init__.py:9-16) and
    ``lib.loss.loss_dict`` (lib/loss.py:739-750).  No reference file is edited; the modules are looked up
    in ``sys.modules`` unless passed explicitly.  Returns the names that were replaced."""
    import sys
    nf = network_factory_module or sys.modules.get("lib.network_factory")
    ls = loss_module or sys.modules.get("lib.loss")
    if nf is None or ls is None:
        raise RuntimeError("deepatlas_b200.install(): import lib.network_factory and lib.loss from the reference "
                           "tree first (or pass the modules explicitly)")
    replaced = []
    for k, v in network_dic.items():
        nf.network_dic[k] = v
        replaced.append("network:" + k)
    for k, v in loss_dict.items():
        ls.loss_dict[k] = v
        replaced.append("loss:" + k)
    return replaced

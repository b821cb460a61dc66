"""Import the REAL reference modules from /root/reference (build container only).  TEST INFRASTRUCTURE.

The reference needs SimpleITK / tensorboardX / matplotlib / skimage, none of which exist in this
image; the hot-path modules only touch three SimpleITK constants at import time, so in-memory stub
modules are enough.  Nothing is copied: the reference is imported where it
lies.  On the GPU box /root/reference does not exist and ``available()`` is False; the golden
fixtures under tests/golden/ (made by oracle/make_golden.py with this loader) stand in for it.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DEEPATLAS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "loss.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


class _Writer:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, *a, **k):
        self.scalars.append(a)

    def add_image(self, *a, **k):
        pass

    def close(self):
        pass


def load(with_models: bool = False):
    """Returns a namespace with the reference's registries and hot-path symbols."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub("SimpleITK", sitkLinear=2, sitkBSpline=3, sitkNearestNeighbor=1)
    if with_models:
        _stub("tensorboardX", SummaryWriter=_Writer)
        for n in ("matplotlib", "matplotlib.backends", "matplotlib.pyplot"):
            _stub(n)
        _stub("matplotlib.backends.backend_agg", FigureCanvasAgg=object)
        _stub("matplotlib.figure", Figure=object)
        _stub("skimage", color=None)
    ns = types.SimpleNamespace()
    import lib.network_factory as nf
    import lib.loss as loss
    import lib.utils as utils
    import lib.transforms as transforms
    ns.network_factory, ns.loss, ns.utils, ns.transforms = nf, loss, utils, transforms
    ns.get_network, ns.network_dic = nf.get_network, nf.network_dic
    ns.get_loss_function, ns.loss_dict = loss.get_loss_function, loss.loss_dict
    if with_models:
        import models.segmentation as seg
        ns.segmentation = seg
    return ns

"""Swap the B200 modules into the reference package so its scripts pick them up unchanged.

    import dvq.patch; dvq.patch.install()     # before `import network.gen_net`
    from network.gen_net import GenNet         # now built from dvq.VectorQuantizer / dvq.PointNetEncoder

``network/VQVAE.py:6`` does ``from network.vqvae.quantizer import VectorQuantizer`` and
``network/gen_net.py:8`` / ``network/DVQVAE.py:8`` do ``from network.pointnet_encoder import
PointNetEncoder`` at import time, so the attributes must be replaced before those imports;
modules that were already imported are re-pointed as well.
"""
from __future__ import annotations

import importlib
import sys


def install(reference_root: str | None = None):
    from .pointnet import PointNetEncoder, STN3d
    from .quantizer import VectorQuantizer
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    q = importlib.import_module("network.vqvae.quantizer")
    q.VectorQuantizer = VectorQuantizer
    p = importlib.import_module("network.pointnet_encoder")
    p.PointNetEncoder = PointNetEncoder
    p.STN3d = STN3d
    for name, attr, obj in (("network.VQVAE", "VectorQuantizer", VectorQuantizer),
                            ("network.gen_net", "PointNetEncoder", PointNetEncoder),
                            ("network.DVQVAE", "PointNetEncoder", PointNetEncoder)):
        mod = sys.modules.get(name)
        if mod is not None:
            setattr(mod, attr, obj)
    return VectorQuantizer, PointNetEncoder

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


def install(reference_root: str | None = None, patch_training_model: bool = False):
    """Swap the modules into the reference's INFERENCE graph (``network.gen_net``, ``network.VQVAE``).
    ``network.DVQVAE`` (whose forward is the training graph) keeps the reference's PointNetEncoder unless
    ``patch_training_model=True``: dvq.PointNetEncoder is inference-only (eval-mode BN folded, no autograd)."""
    from .pointnet import PointNetEncoder, STN3d
    from .quantizer import VectorQuantizer
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    q = importlib.import_module("network.vqvae.quantizer")
    q.VectorQuantizer = VectorQuantizer
    # network/gen_net.py:8 and network/DVQVAE.py:8 both do `from network.pointnet_encoder import PointNetEncoder` at
    # import time, so the attribute of network.pointnet_encoder cannot be swapped globally without also reaching the
    # training model: import the consumers first, then re-point only the ones asked for
    importlib.import_module("network.pointnet_encoder")
    importlib.import_module("network.VQVAE").VectorQuantizer = VectorQuantizer
    gen = importlib.import_module("network.gen_net")
    gen.PointNetEncoder = PointNetEncoder
    if patch_training_model:
        p = sys.modules["network.pointnet_encoder"]
        p.PointNetEncoder = PointNetEncoder
        p.STN3d = STN3d
        importlib.import_module("network.DVQVAE").PointNetEncoder = PointNetEncoder
    return VectorQuantizer, PointNetEncoder

"""Shared case table: regenerates the seeded inputs of oracle/gen_golden.py
without the reference (numpy legacy RandomState is bit-stable)."""
import hashlib
import os

import numpy as np

from oracle import vq_oracle as vo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

VQ_CASES = {
    "vq_k512_d64": (1024, 512, 64, 1.0, 0.25, "default", None),
    "vq_k128_d256": (512, 128, 256, 1.0, 0.25, "default", None),
    "vq_k128_d1024": (256, 128, 1024, 0.0, 2.0, "default", None),
    "vq_variant_b": (1024, 512, 64, 1.0, 0.25, "variant_b", None),
    "vq_dupes": (777, 64, 16, 1.0, 0.25, "dupes", None),
    "vq_3d_view": (4 * 9 * 5, 96, 24, 0.5, 0.75, "default", (4, 9, 5, 24)),
    "vq_ragged": (333, 200, 40, 1.0, 0.25, "default", None),
    "vq_one_row": (1, 128, 256, 1.0, 0.25, "default", None),
}
PN_CASES = ["pointnet_c4_p3000", "pointnet_c3_p778", "pointnet_c4_p100", "pointnet_c4_p129"]


def vq_inputs(name):
    n, k, d, al, beta, kind, shape = VQ_CASES[name]
    seed = int(hashlib.sha256(name.encode()).hexdigest()[:6], 16)
    if kind == "variant_b":
        z, E = vo.variant_b(n, k, d, seed)
    else:
        E = vo.default_codebook(k, d, seed)
        z = vo.normal_latents(n, d, seed + 1)
        if kind == "dupes":
            E[k // 2:] = E[:k // 2]
    if shape is not None:
        z = z.reshape(shape)
    return z, E, al, beta


def load_gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def rel_err(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-30)

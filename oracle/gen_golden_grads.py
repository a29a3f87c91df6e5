"""Generate tests/golden/vq_grads_*.npz: gradients of the REAL reference's autograd graph
(network/vqvae/quantizer.py:36-43,56-60 — the `.detach()` placement decides which term reaches z
and which the codebook) for the backward of dvq.VectorQuantizer.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden_grads.py
Objective: L = 3 * loss + sum(z_q ** 2) + 0.5 * sum(z_q * w), w a fixed seeded tensor — exercises both the
loss path (al / beta terms) and the straight-through path (d z_q / d z = 1, d z_q / d E = 0)."""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

import network.vqvae.quantizer as refq  # noqa: E402
from _cases import vq_inputs  # noqa: E402

refq.device = torch.device("cpu")
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = ["vq_ragged", "vq_3d_view", "vq_k512_d64"]


def weights_for(name, shape):
    seed = int(hashlib.sha256(("grad" + name).encode()).hexdigest()[:6], 16)
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32)


def main():
    for name in CASES:
        z, E, al, beta = vq_inputs(name)
        m = refq.VectorQuantizer(E.shape[0], E.shape[1], beta, al)
        with torch.no_grad():
            m.embedding.weight.copy_(torch.from_numpy(E))
        zt = torch.from_numpy(z).clone().requires_grad_(True)
        w = torch.from_numpy(weights_for(name, z.shape))
        loss, zq, ppl, enc, idx = m(zt, True)
        obj = 3.0 * loss + (zq * zq).sum() + 0.5 * (zq * w).sum()
        obj.backward()
        np.savez_compressed(os.path.join(GOLD, "vq_grads_%s.npz" % name), dz=zt.grad.numpy(), dE=m.embedding.weight.grad.numpy(),
                            idx=idx.numpy().astype(np.int32), obj=np.float32(obj.item()))
        print(name, float(obj.item()), float(zt.grad.abs().max()), float(m.embedding.weight.grad.abs().max()))


if __name__ == "__main__":
    main()

"""Generate tests/golden/* by running the REAL reference on CPU.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py
The reference module global ``device`` (network/vqvae/quantizer.py:7) is
re-pointed to CPU — the only modification; every tensor op is the reference's.
Inputs come from numpy's legacy RandomState (oracle/vq_oracle.py,
oracle/pointnet_oracle.py) so tests can regenerate them without the reference;
only the reference's OUTPUTS are stored.  Config 1 additionally follows the
BASELINE.md recipe (torch.manual_seed(0); construct; z = torch.randn).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import network.vqvae.quantizer as refq  # noqa: E402
from network.pointnet_encoder import PointNetEncoder as RefPointNet  # noqa: E402
from network.VQVAE import VQVAE as RefVQVAE  # noqa: E402

from oracle import pointnet_oracle as po  # noqa: E402
from oracle import ref_port_torch as port  # noqa: E402
from oracle import vq_oracle as vo  # noqa: E402

refq.device = torch.device("cpu")
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)

VQ_CASES = {
    # name: (N, K, D, al, beta, kind, z_shape)
    "vq_k512_d64": (1024, 512, 64, 1.0, 0.25, "default", None),
    "vq_k128_d256": (512, 128, 256, 1.0, 0.25, "default", None),       # part codebooks, DVQVAE.py:23-28
    "vq_k128_d1024": (256, 128, 1024, 0.0, 2.0, "default", None),      # object codebook, DVQVAE.py:29
    "vq_variant_b": (1024, 512, 64, 1.0, 0.25, "variant_b", None),
    "vq_dupes": (777, 64, 16, 1.0, 0.25, "dupes", None),               # exact ties -> lowest index
    "vq_3d_view": (4 * 9 * 5, 96, 24, 0.5, 0.75, "default", (4, 9, 5, 24)),
    "vq_ragged": (333, 200, 40, 1.0, 0.25, "default", None),           # K, N not multiples of any tile
    "vq_one_row": (1, 128, 256, 1.0, 0.25, "default", None),
}


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


def run_ref_vq(z, E, al, beta):
    k, d = E.shape
    m = refq.VectorQuantizer(k, d, beta, al)
    with torch.no_grad():
        m.embedding.weight.copy_(torch.from_numpy(E))
        zt = torch.from_numpy(z)
        loss, zq_t, ppl, enc, idx_t = m(zt, True)
        idx_i, zq_i = m(zt, False)
    return m, dict(
        idx_train=idx_t.numpy().astype(np.int32), idx_infer=idx_i.numpy().astype(np.int32),
        zq_train=zq_t.numpy(), zq_infer=zq_i.numpy(),
        loss=np.float32(loss.item()), perplexity=np.float32(ppl.item()),
        hist=enc.sum(0).numpy().astype(np.int64))


def main():
    port_report = {}
    for name in VQ_CASES:
        z, E, al, beta = vq_inputs(name)
        m, out = run_ref_vq(z, E, al, beta)
        assert out["idx_train"].shape == (z.size // E.shape[1], 1)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        # the torch port must be bit-identical to the reference (same ops, same library)
        pl, pq, pp, ph, pi = port.quantize_rows(torch.from_numpy(z), torch.from_numpy(E), al, beta, True)
        ii, iq = port.quantize_rows(torch.from_numpy(z), torch.from_numpy(E), al, beta, False)
        port_report[name] = dict(
            idx_equal=bool((pi.numpy() == out["idx_train"]).all() and (ii.numpy() == out["idx_infer"]).all()),
            zq_bits_equal=bool(np.array_equal(pq.numpy().view(np.uint32), out["zq_train"].view(np.uint32))
                               and np.array_equal(iq.numpy().view(np.uint32), out["zq_infer"].view(np.uint32))),
            loss_equal=bool(np.float32(pl.item()) == out["loss"]), ppl_equal=bool(np.float32(pp.item()) == out["perplexity"]))
        print(name, float(out["loss"]), float(out["perplexity"]), port_report[name])

    # get_emb (quantizer.py:68-75) through the VQVAE wrapper (VQVAE.py:51-53), one index at a time (B=1 semantics)
    z, E, al, beta = vq_inputs("vq_k128_d256")
    w = RefVQVAE(128, 32, 2, 128, 256, 0.25, a=1)
    with torch.no_grad():
        w.vector_quantization.embedding.weight.copy_(torch.from_numpy(E))
        picks = np.array([0, 5, 127, 64], dtype=np.int64)
        embs = np.stack([w.get_embbeding(torch.tensor([p]), 256).numpy()[0] for p in picks])
        # wrapper forward / inference return arity (VQVAE.py:29-50)
        l3, q3, p3 = w(torch.from_numpy(z))
        i2, q2 = w.inference(torch.from_numpy(z))
    np.savez_compressed(os.path.join(GOLD, "vq_get_emb.npz"), picks=picks, embs=embs,
                        wrapper_loss=np.float32(l3.item()), wrapper_ppl=np.float32(p3.item()),
                        wrapper_idx=i2.numpy().astype(np.int32))

    # BASELINE config 1, BASELINE.md recipe (torch RNG): full 65 536 rows
    torch.manual_seed(0)
    m = refq.VectorQuantizer(512, 64, 0.25, 1)
    zt = torch.randn(65536, 64)
    with torch.no_grad():
        loss, zq, ppl, enc, idx = m(zt, True)
        idx_i, zq_i = m(zt, False)
    assert (idx == idx_i).all()
    np.savez_compressed(
        os.path.join(GOLD, "vq_config1_full.npz"),
        idx=idx.numpy().astype(np.uint16).reshape(-1), loss=np.float32(loss.item()), perplexity=np.float32(ppl.item()),
        zq_train_sha256=hashlib.sha256(zq.numpy().tobytes()).hexdigest(),
        codebook_sha256=hashlib.sha256(m.embedding.weight.detach().numpy().tobytes()).hexdigest(),
        z_sha256=hashlib.sha256(zt.numpy().tobytes()).hexdigest())
    print("config1", loss.item(), ppl.item())

    # PointNet encoder (pointnet_encoder.py:125-169), eval mode
    pn_cases = {"pointnet_c4_p3000": (2, 4, 3000, 101), "pointnet_c3_p778": (3, 3, 778, 202),
                "pointnet_c4_p100": (1, 4, 100, 303), "pointnet_c4_p129": (5, 4, 129, 404)}
    for name, (b, c, p, seed) in pn_cases.items():
        sd = po.make_state(seed, c)
        x = po.make_cloud(seed + 1, b, c, p)
        net = RefPointNet(global_feat=True, feature_transform=False, channel=c).eval()
        net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        with torch.no_grad():
            feat, trans, tf = net(torch.from_numpy(x))
        assert tf is None
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), feat=feat.numpy(), trans=trans.numpy(),
                            meta=np.array([b, c, p, seed], dtype=np.int64))
        pf, pt, _ = port.pointnet_eval(torch.from_numpy(x), {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        port_report[name] = dict(feat_max_abs_diff=float((pf - feat).abs().max()), trans_max_abs_diff=float((pt - trans).abs().max()))
        print(name, feat.abs().max().item(), port_report[name])

    with open(os.path.join(GOLD, "port_vs_reference.json"), "w") as f:
        json.dump(port_report, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

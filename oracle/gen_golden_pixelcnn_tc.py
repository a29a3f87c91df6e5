"""Golden vectors for the kernel-backed PixelCNN sampler (precision "fp16_tc") — TEST INFRASTRUCTURE ONLY (runs where
/root/reference exists).  A model wide enough for the tcgen05 GEMM tiles (dim 256, 256 input classes) is built from the
REAL reference class under a fixed seed; only the logits of every sampling step (and a digest of the state dict) are
stored: the test rebuilds the same weights from the seed — dvq's GatedPixelCNN mirrors the reference's constructor, so the
RNG stream is consumed identically — and verifies the digest before comparing.  Writes tests/golden/pixelcnn_tc.npz."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from network.pixelcnn.models import GatedPixelCNN  # noqa: E402

CFG = dict(input_dim=256, dim=256, n_layers=3, n_classes=8)
SEED, B = 7, 5


def build(cls):
    torch.manual_seed(SEED)
    m = cls(**CFG).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.copy_(0.1 * torch.randn_like(p))
    return m


def digest(m):
    h = hashlib.sha256()
    for k, v in sorted(m.state_dict().items()):
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    ref = build(GatedPixelCNN)
    g = torch.Generator().manual_seed(99)
    x_full = torch.randint(0, CFG["input_dim"], (B, 3, 3), generator=g)
    label = torch.randint(0, CFG["n_classes"], (B,), generator=g)
    out = {"cfg": np.array([CFG["input_dim"], CFG["dim"], CFG["n_layers"], CFG["n_classes"]]), "x_full": x_full.numpy(), "label": label.numpy()}
    with torch.no_grad():
        x = torch.zeros_like(x_full)
        for i in range(3):
            for j in range(3):
                out["logits_%d%d" % (i, j)] = ref(x, label)[:, :, i, j].numpy().copy()
                x[:, i, j] = x_full[:, i, j]
    out["sd_sha256"] = np.array(digest(ref))       # after the forwards: layer 0 carries its mask (make_causal)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pixelcnn_tc.npz"), **out)
    print("wrote pixelcnn_tc.npz", float(np.abs(out["logits_22"]).max()), out["sd_sha256"])


if __name__ == "__main__":
    main()

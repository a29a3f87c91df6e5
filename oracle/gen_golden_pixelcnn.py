"""Golden vectors for the PixelCNN sampler — TEST INFRASTRUCTURE ONLY (runs where /root/reference exists).

Imports the real reference ``network.pixelcnn.models.GatedPixelCNN`` (CPU), builds a small random-init
model, and records for a fixed index grid the logits the reference's *full* forward gives at every
position (i, j) when the grid holds the true indices before (i, j) in raster order and zeros after — i.e.
exactly what ``generate`` (models.py:187-197) evaluates at that step — plus one seeded ``generate`` sample.
Writes tests/golden/pixelcnn_small.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from network.pixelcnn.models import GatedPixelCNN  # noqa: E402


def main():
    torch.manual_seed(7)
    cfg = dict(input_dim=16, dim=8, n_layers=4, n_classes=6)
    ref = GatedPixelCNN(**cfg).eval()
    # the reference initialises conv biases to 0 and leaves embeddings N(0,1); perturb biases so they matter
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if n.endswith("bias"):
                p.copy_(0.1 * torch.randn_like(p))
    B = 5
    x_full = torch.randint(0, cfg["input_dim"], (B, 3, 3))
    label = torch.randint(0, cfg["n_classes"], (B,))
    out = {"cfg": np.array([cfg["input_dim"], cfg["dim"], cfg["n_layers"], cfg["n_classes"]]), "x_full": x_full.numpy(), "label": label.numpy()}
    with torch.no_grad():
        x = torch.zeros_like(x_full)
        for i in range(3):
            for j in range(3):
                logits = ref(x, label)[:, :, i, j]
                out["logits_%d%d" % (i, j)] = logits.numpy().copy()
                x[:, i, j] = x_full[:, i, j]
        out["full_logits"] = ref(x_full, label).numpy()
        torch.manual_seed(123)
        out["sample_seed123"] = ref.generate(x_full, label, shape=(3, 3), batch_size=B).numpy()
    for k, v in ref.state_dict().items():     # after the forwards: layer 0 carries its mask (make_causal)
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pixelcnn_small.npz"), **out)
    print("wrote pixelcnn_small.npz", {k: v.shape for k, v in out.items() if not k.startswith("sd.")})


if __name__ == "__main__":
    main()

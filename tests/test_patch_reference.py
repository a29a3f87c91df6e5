"""CPU, build container only (skipped where /root/reference is absent): dvq.patch.install() makes the
reference's own GenNet / DVQVAE constructors build from the B200 modules, and the resulting
state_dict has exactly the reference's keys and shapes (checkpoints load unchanged)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, json
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/d-vqvae_b200"); sys.path.insert(0, %(ref)r)
patched = %(patched)s
if patched:
    import dvq.patch
    dvq.patch.install()
from network.gen_net import GenNet
import torch
torch.manual_seed(0)
net = GenNet()
sd = {k: list(v.shape) for k, v in net.state_dict().items()}
kinds = {n: type(m).__module__ for n, m in net.named_modules() if n in ("obj_encoder_type", "vqvae6.vector_quantization", "recon_encoder")}
print(json.dumps({"sd": sd, "kinds": kinds}))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_patch_install_swaps_modules_and_keeps_state_dict():
    import json
    outs = {}
    for patched in (False, True):
        r = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, ref=REF, patched=patched)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[patched] = json.loads(r.stdout.strip().splitlines()[-1])
    assert outs[False]["sd"] == outs[True]["sd"]                      # same keys, same shapes
    assert all(v.startswith("network.") for v in outs[False]["kinds"].values())
    assert all(v.startswith("dvq.") for v in outs[True]["kinds"].values()), outs[True]["kinds"]

"""Golden vectors of the REAL reference's generation graph (network/gen_net.py:78-125) — TEST INFRASTRUCTURE ONLY
(runs in the build container where /root/reference exists).

``GenNet.gen`` itself is executed: its source is taken with ``inspect`` and run with the single token
``device='cuda'`` -> ``device='cpu'`` (the same kind of patch as ``quantizer.device``); every sub-module is the
reference's own class.  The three things a random-init, asset-free run needs are supplied from outside, identically on
both sides: (1) the weights — a ``dvq.GraspGenerator`` built under ``torch.manual_seed(0)`` (the test rebuilds it from
the seed and checks the digest) copied into the reference ``GenNet`` (same parameter names); (2) the prior — a stub
``generate`` returning fixed code grids (a random-init reference PixelCNN samples classes >= 128 that ``get_emb`` cannot
index); (3) the MANO layer — ``dvq.grasp.LinearHandStub`` behind the ``rh_mano(...).vertices`` call shape.
Reference semantics are B = 1 (gen_net.py:88-89), so the objects go through one at a time.
Writes tests/golden/gennet_stages.npz."""
import hashlib
import inspect
import os
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
sys.path.insert(0, "/root/reference")

import network.vqvae.quantizer as refq  # noqa: E402
from network.gen_net import GenNet  # noqa: E402

import dvq  # noqa: E402
from dvq.grasp import LinearHandStub  # noqa: E402
from oracle import pointnet_oracle as po  # noqa: E402

refq.device = torch.device("cpu")
N_OBJ, P = 3, 3000


def digest(sd):
    h = hashlib.sha256()
    for k, v in sorted(sd.items()):
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


class _Verts:
    def __init__(self, v):
        self.vertices = v


class _ManoStub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.stub = LinearHandStub()

    def forward(self, betas, global_orient, hand_pose, transl):
        return _Verts(self.stub(betas, hand_pose))


class _PriorStub(torch.nn.Module):
    def __init__(self, grid):
        super().__init__()
        self.grid = grid

    def generate(self, idx6, label, shape=(3, 3), batch_size=1):
        return self.grid


def main():
    torch.manual_seed(0)
    ours = dvq.GraspGenerator()
    sd = ours.state_dict()
    ref = GenNet()
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("GatedPixelCNN.") for k in missing), (unexpected, [k for k in missing if not k.startswith("GatedPixelCNN.")][:5])
    ref.eval()
    ref.rh_mano = _ManoStub()
    src = textwrap.dedent(inspect.getsource(GenNet.gen)).replace("device='cuda'", "device='cpu'")
    assert "device='cpu'" in src
    ns = {"torch": torch}
    exec(compile(src, "<network/gen_net.py:gen, device -> cpu>", "exec"), ns)
    gen_cpu = types.MethodType(ns["gen"], ref)
    rs = np.random.RandomState(77)
    codes = rs.randint(0, 128, size=(N_OBJ, 3, 3)).astype(np.int64)
    clouds = po.make_cloud(501, N_OBJ, 4, P)
    recon, pos, feat_t, idx6 = [], [], [], []
    with torch.no_grad():
        for i in range(N_OBJ):
            ref.GatedPixelCNN = _PriorStub(torch.from_numpy(codes[i:i + 1]))
            obj = torch.from_numpy(clouds[i:i + 1])
            r, p = gen_cpu(obj)
            recon.append(r.numpy()); pos.append(p.numpy())
            f, _, _ = ref.obj_encoder_type(obj)
            feat_t.append(f.numpy())
            idx6.append(ref.vqvae6.inference(f)[0].numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gennet_stages.npz"), codes=codes, recon=np.concatenate(recon), recon_pos=np.concatenate(pos),
                        feat_type=np.concatenate(feat_t), idx6=np.concatenate(idx6).reshape(-1), sd_sha256=np.array(digest(sd)), cloud_seed=np.array(501))
    print("wrote gennet_stages.npz", np.concatenate(recon).shape, float(np.abs(np.concatenate(recon)).max()), digest(sd)[:16])


if __name__ == "__main__":
    main()

"""PixelCNN prior: the row-cached sampler against golden vectors of the real reference
(tests/golden/pixelcnn_small.npz, made by oracle/gen_golden_pixelcnn.py) and against its own full forward.
CPU tests (the sampler is torch code over library GEMMs); the GPU variant checks the same on cuda:0."""
import os

import numpy as np
import pytest
import torch

import dvq
from dvq.pixelcnn import GatedPixelCNN

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pixelcnn_small.npz")


def _load():
    g = np.load(GOLD)
    input_dim, dim, n_layers, n_classes = (int(v) for v in g["cfg"])
    m = GatedPixelCNN(input_dim, dim, n_layers, n_classes).eval()
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    assert set(sd) == set(m.state_dict())                       # same parameter names and shapes as the reference
    m.load_state_dict(sd)
    return g, m


def test_state_dict_layout_matches_reference_and_full_forward():
    g, m = _load()
    x, label = torch.from_numpy(g["x_full"]), torch.from_numpy(g["label"])
    with torch.no_grad():
        out = m(x, label)
    assert np.allclose(out.numpy(), g["full_logits"], rtol=1e-5, atol=1e-5)


def test_cached_sampler_logits_equal_reference_full_forward_at_every_step():
    g, m = _load()
    x_full, label = torch.from_numpy(g["x_full"]), torch.from_numpy(g["label"])
    x, logits = m.generate(x_full, label, shape=(3, 3), batch_size=x_full.shape[0], forced=x_full, return_logits=True)
    assert torch.equal(x, x_full)
    for n, (i, j) in enumerate((i, j) for i in range(3) for j in range(3)):
        ref = g["logits_%d%d" % (i, j)]
        assert np.allclose(logits[n].numpy(), ref, rtol=1e-5, atol=1e-5), (i, j, np.abs(logits[n].numpy() - ref).max())


def test_sampling_reproduces_reference_generate_under_the_same_seed():
    g, m = _load()
    x_full, label = torch.from_numpy(g["x_full"]), torch.from_numpy(g["label"])
    torch.manual_seed(123)
    x = m.generate(x_full, label, shape=(3, 3), batch_size=x_full.shape[0])
    assert np.array_equal(x.numpy(), g["sample_seed123"])


@pytest.mark.parametrize("cfg", [(32, 16, 6, 10), (64, 24, 15, 128)])
def test_cached_sampler_equals_own_full_forward(cfg):
    torch.manual_seed(11)
    m = GatedPixelCNN(*cfg).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.copy_(0.05 * torch.randn_like(p))
    B = 7
    x_full = torch.randint(0, cfg[0], (B, 3, 3))
    label = torch.randint(0, cfg[3], (B,))
    _, logits = m.generate(None, label, shape=(3, 3), batch_size=B, forced=x_full, return_logits=True)
    x = torch.zeros_like(x_full)
    n = 0
    with torch.no_grad():
        for i in range(3):
            for j in range(3):
                ref = m(x, label)[:, :, i, j]
                assert torch.allclose(logits[n], ref, rtol=2e-5, atol=2e-5), (i, j)
                x[:, i, j] = x_full[:, i, j]
                n += 1


def test_n_valid_restricts_the_sampled_classes():
    torch.manual_seed(3)
    m = GatedPixelCNN(64, 16, 3, 8).eval()
    x = m.generate(None, torch.zeros(9, dtype=torch.int64), batch_size=9, n_valid=5)
    assert int(x.max()) < 5 and x.shape == (9, 3, 3)


@pytest.mark.gpu
def test_cached_sampler_on_gpu_and_as_grasp_prior():
    g, m = _load()
    m = m.cuda()
    x_full, label = torch.from_numpy(g["x_full"]).cuda(), torch.from_numpy(g["label"]).cuda()
    _, logits = m.generate(x_full, label, shape=(3, 3), batch_size=x_full.shape[0], forced=x_full, return_logits=True)
    for n, (i, j) in enumerate((i, j) for i in range(3) for j in range(3)):
        assert np.allclose(logits[n].cpu().numpy(), g["logits_%d%d" % (i, j)], rtol=1e-4, atol=1e-4)
    # as the prior of the grasp pipeline (gen_net.py:92-100): random-init PixelCNN restricted to the 128 codebook rows
    from dvq.grasp import GraspGenerator, pixelcnn_prior
    torch.manual_seed(0)
    prior_net = GatedPixelCNN(512, 64, 3, 128).cuda().eval()
    gen = GraspGenerator(prior=pixelcnn_prior(prior_net, n_valid=128)).cuda().eval()
    obj = torch.randn(8, 4, 256, device="cuda") * 0.1
    recon, pos = gen.gen(obj)
    assert recon.shape == (8, 55) and pos.shape == (8, 6) and torch.isfinite(recon).all()

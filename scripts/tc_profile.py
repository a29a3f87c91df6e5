"""Minimal config-2 driver for ncu captures: a few train-path forwards of the VQ module."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch, dvq
N, K, D = int(os.environ.get("N", 4194304)), int(os.environ.get("K", 512)), int(os.environ.get("D", 64))
g = torch.Generator(device="cuda").manual_seed(2000)
vq = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda()
vq.onehot_limit_bytes = 0
with torch.no_grad():
    vq.embedding.weight.copy_((torch.rand(K, D, device="cuda", generator=g) * 2 - 1) / K)
    z = torch.randn(N, D, device="cuda", generator=g)
    for _ in range(int(os.environ.get("ITERS", 4))):
        out = vq(z, True)
torch.cuda.synchronize()
print("loss", out[0].item(), "ppl", out[2].item(), "counters", vq.last_counters(N))

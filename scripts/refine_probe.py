"""Undecided / overflow row counts and refine time of one streamed shape (N, K, D from the environment)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch, dvq
from dvq import _cabi
N, K, D = int(os.environ.get("N", 2097152)), int(os.environ.get("K", 16384)), int(os.environ.get("D", 256))
g = torch.Generator(device="cuda").manual_seed(2000)
vq = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda(); vq.onehot_limit_bytes = 0
with torch.no_grad():
    vq.embedding.weight.copy_((torch.rand(K, D, device="cuda", generator=g) * 2 - 1) / K)
    z = torch.randn(N, D, device="cuda", generator=g)
    for _ in range(2): out = vq(z, True)
    torch.cuda.synchronize()
    _cabi.lib.dvq_profile_enable(1)
    for _ in range(3): out = vq(z, True)
    torch.cuda.synchronize()
    ms, cnt = _cabi.profile_mean(); _cabi.lib.dvq_profile_enable(0)
o = (C.c_int * 4)()
_cabi.lib.dvq_vq_read_counters(vq._ws.data_ptr(), N, K, D, vq.path, o)
print("N %d K %d D %d  kernel %.3f ms refine %.3f ms  listed %d overflow %d (%.2f %% / %.3f %% of rows)" % (N, K, D, ms[1], ms[2], o[0], o[2], 100.0 * o[0] / N, 100.0 * o[2] / N))

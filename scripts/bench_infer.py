import sys; sys.path.insert(0, "d-vqvae_b200"); sys.path.insert(0, ".")
import torch, dvq
from dvq import _cabi
N, K, D = 4194304, 512, 64
g = torch.Generator(device="cuda").manual_seed(2000)
vq = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda(); vq.onehot_limit_bytes = 0
with torch.no_grad():
    vq.embedding.weight.copy_((torch.rand(K, D, device="cuda", generator=g) * 2 - 1) / K)
    z = torch.randn(N, D, device="cuda", generator=g)
    for train in (True, False, True, False):
        for _ in range(5): out = vq(z, train)
        torch.cuda.synchronize()
        _cabi.lib.dvq_profile_enable(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): out = vq(z, train)
        b.record(); torch.cuda.synchronize()
        ms, cnt = _cabi.profile_mean(); _cabi.lib.dvq_profile_enable(0)
        print("train" if train else "infer", "step_ms %.4f kernel %.4f refine %.4f" % (a.elapsed_time(b) / 20, ms[1], ms[2]))

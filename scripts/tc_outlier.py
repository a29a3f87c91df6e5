"""Per-call timing of the tcgen05 path for configs that showed sporadic slow calls."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch, dvq
from dvq import _cabi

def run(N, K, D, reps=8, path=_cabi.DVQ_PATH_TC):
    g = torch.Generator(device="cuda").manual_seed(0)
    E = (torch.rand(K, D, device="cuda", generator=g) * 2 - 1) / K
    z = torch.randn(N, D, device="cuda", generator=g)
    m = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda(); m.path = path; m.onehot_limit_bytes = 0
    rows = []
    with torch.no_grad():
        m.embedding.weight.copy_(E)
        for i in range(reps):
            torch.cuda.synchronize()
            _cabi.lib.dvq_profile_enable(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            r = m(z, True)
            e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
            ms, cnt = _cabi.profile_mean(); _cabi.lib.dvq_profile_enable(0)
            rows.append((round(e0.elapsed_time(e1), 3), round((t1 - t0) * 1e3, 3), [round(v, 3) for v in ms]))
    print((N, K, D), "event_ms, wall_ms, [norms, main, refine, onehot]")
    for r in rows: print("   ", r)

for cfg in [(65536, 512, 64), (4194304, 512, 64), (100000, 1536, 32), (4194304, 512, 64), (100000, 1536, 32)]:
    run(*cfg)

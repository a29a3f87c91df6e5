"""GPU sweep of the tcgen05 VQ path over (N, K, D): per-stage CUDA-event times and index parity vs the FP32 kernel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch
import dvq
from dvq import _cabi

def one(N, K, D, reps=5):
    g = torch.Generator(device="cuda").manual_seed(1)
    E = (torch.rand(K, D, device="cuda", generator=g) * 2 - 1) / K
    z = torch.randn(N, D, device="cuda", generator=g)
    out = {}
    idx = {}
    for name, path in (("simt", _cabi.DVQ_PATH_SIMT), ("tc", _cabi.DVQ_PATH_TC)):
        m = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda()
        m.path = path; m.onehot_limit_bytes = 0
        with torch.no_grad():
            m.embedding.weight.copy_(E)
            r = m(z, True); torch.cuda.synchronize()
            _cabi.lib.dvq_profile_enable(1)
            for _ in range(reps):
                r = m(z, True)
            torch.cuda.synchronize()
            ms, cnt = _cabi.profile_mean()
            _cabi.lib.dvq_profile_enable(0)
        idx[name] = r[4]
        out[name] = dict(norms=round(ms[0], 4), main=round(ms[1], 4), refine=round(ms[2], 4), counters=m.last_counters(N))
    out["mismatch"] = int((idx["simt"] != idx["tc"]).sum())
    out["tflops_main"] = round(2.0 * N * K * D / (out["tc"]["main"] * 1e-3) / 1e12, 1)
    return out

if __name__ == "__main__":
    cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(100000, 1536, 32), (100000, 1536, 64), (100000, 1024, 32), (100000, 2048, 32), (100000, 768, 32)]
    rep = {}
    for c in cfgs:
        try:
            rep[str(c)] = one(*c)
        except Exception as e:
            rep[str(c)] = "EXC: %r" % (e,)
        print(c, rep[str(c)], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rep, open("gpurun_out/tc_sweep.json", "w"), indent=1, default=str)

"""GPU debug driver for the tcgen05 VQ path: small and full-size runs vs the FP32 kernel."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import numpy as np, torch
import dvq
from dvq import _cabi
from oracle import vq_oracle as vo

def run(N, K, D, variant="default", train=True, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if variant == "default":
        E = (torch.rand(K, D, device="cuda", generator=g) * 2 - 1) / K
        z = torch.randn(N, D, device="cuda", generator=g)
    else:
        E = torch.randn(K, D, device="cuda", generator=g)
        z = E[torch.randint(0, K, (N,), device="cuda", generator=g)] + 0.1 * torch.randn(N, D, device="cuda", generator=g)
    out = {}
    res = {}
    for name, path in (("simt", _cabi.DVQ_PATH_SIMT), ("tc", _cabi.DVQ_PATH_TC)):
        m = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda()
        m.path = path; m.onehot_limit_bytes = 0
        with torch.no_grad():
            m.embedding.weight.copy_(E)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = m(z, True) if train else m(z, False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            # timed second run
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(3):
                r = m(z, True) if train else m(z, False)
            ev1.record(); torch.cuda.synchronize()
        res[name] = r
        out[name + "_ms"] = ev0.elapsed_time(ev1) / 3
        out[name + "_counters"] = m.last_counters(N)
    if train:
        ls, zs, ps, _, is_ = res["simt"]; lt, zt, pt, _, it = res["tc"]
        out["loss"] = (ls.item(), lt.item()); out["ppl"] = (ps.item(), pt.item())
    else:
        is_, zs = res["simt"]; it, zt = res["tc"]
    mism = (is_ != it).view(-1)
    out["idx_mismatch"] = int(mism.sum())
    out["zq_equal_where_idx_equal"] = bool(torch.equal(zs[~mism], zt[~mism]))
    if out["idx_mismatch"]:
        rows = mism.nonzero().view(-1)[:4096].cpu().numpy()
        zz = z[rows].cpu().numpy(); En = E.cpu().numpy()
        n_mis, n_bad, worst = vo.allowed_index_mismatch(zz, En, it.view(-1)[rows].cpu().numpy(), is_.view(-1)[rows].cpu().numpy())
        out["band_check"] = (n_mis, n_bad, worst)
    return out

if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    rep = {}
    for cfg in [(128, 32, 16, "default"), (1000, 512, 64, "default"), (65536, 512, 64, "default"), (65536, 512, 64, "variant_b"),
                (4194304, 512, 64, "default"), (4194304, 512, 64, "variant_b"), (300000, 256, 32, "default"), (100000, 1024, 16, "default"),
                (262144, 4096, 64, "default"), (262144, 4096, 64, "variant_b"), (65536, 16384, 64, "default"), (100000, 1536, 32, "default"), (50000, 800, 64, "default")]:
        try:
            rep[str(cfg)] = run(*cfg)
        except Exception as e:
            rep[str(cfg)] = "EXC: %r" % (e,)
            print(cfg, rep[str(cfg)]); break
        print(cfg, rep[str(cfg)], flush=True)
    json.dump(rep, open("gpurun_out/tc_debug.json", "w"), indent=1, default=str)

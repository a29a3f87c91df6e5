"""A/B driver for kernel variants at BASELINE config 2 (or N/K/D from the environment): per-stage CUDA-event
times of dvq_vq_forward (library-side events), whole-step time, undecided rows, and the index / z_q agreement
with the all-FP32 kernel of the same library.  One line of JSON per run; DVQ_LIB selects the variant build."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch, dvq
from dvq import _cabi
N, K, D = int(os.environ.get("N", 4194304)), int(os.environ.get("K", 512)), int(os.environ.get("D", 64))
STEPS = int(os.environ.get("STEPS", 30))
g = torch.Generator(device="cuda").manual_seed(2000)
vq = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda()
vq.onehot_limit_bytes = 0
with torch.no_grad():
    vq.embedding.weight.copy_((torch.rand(K, D, device="cuda", generator=g) * 2 - 1) / K)
    z = torch.randn(N, D, device="cuda", generator=g)
    vq.path = _cabi.DVQ_PATH_SIMT
    ref = vq(z, True)
    vq.path = _cabi.DVQ_PATH_AUTO
    for _ in range(5):
        out = vq(z, True)
    torch.cuda.synchronize()
    _cabi.lib.dvq_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        out = vq(z, True)
    e1.record()
    torch.cuda.synchronize()
    ms, cnt = _cabi.profile_mean()
    _cabi.lib.dvq_profile_enable(0)
mism = int((out[4] != ref[4]).sum())
same = (out[4] == ref[4]).view(-1)
zq_ok = bool(torch.equal(out[1][same], ref[1][same]))
print(json.dumps({"tag": os.environ.get("TAG", ""), "N": N, "K": K, "D": D, "step_ms": round(e0.elapsed_time(e1) / STEPS, 4),
                  "prep_ms": round(ms[0], 4), "kernel_ms": round(ms[1], 4), "refine_ms": round(ms[2], 4),
                  "counters": vq.last_counters(N), "idx_mismatch_vs_simt": mism, "zq_equal": zq_ok,
                  "loss": out[0].item(), "loss_simt": ref[0].item(), "ppl": out[2].item(), "ppl_simt": ref[2].item()}))

"""GatedPixelCNN(512, 512, 15) sampler, batch B: the reference algorithm (one full forward per position), the row-cached
sampler on torch GEMMs (tf32) and on the repo's tcgen05 GEMM kernel (fp16_tc).  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch
from dvq.pixelcnn import GatedPixelCNN
from dvq import _cabi
B = int(os.environ.get("B", 4096))
dev = torch.device("cuda")
torch.manual_seed(1)
m = GatedPixelCNN(512, 512, 15).to(dev).eval().requires_grad_(False)
label = torch.randint(0, 128, (B,), device=dev)


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


res = {"B": B}
for prec in os.environ.get("PRECS", "tf32,fp16_tc").split(","):
    m.precision = prec
    l0 = _cabi.launch_count()
    res["sampler_%s_ms" % prec] = timed(lambda: m.generate(None, label, batch_size=B, n_valid=128))
    res["launches_%s" % prec] = (_cabi.launch_count() - l0) // 4
if m._tc_sampler is not None:
    m._tc_sampler.check()
# agreement of the two backends on forced indices
g = torch.Generator(device=dev).manual_seed(3)
xf = torch.randint(0, 128, (B, 3, 3), device=dev, generator=g)
m.precision = "tf32"; _, la = m.generate(None, label, batch_size=B, forced=xf, return_logits=True)
m.precision = "fp16_tc"; _, lb = m.generate(None, label, batch_size=B, forced=xf, return_logits=True)
res["max_rel_logit_diff_vs_tf32"] = max(float((a - b).abs().max() / a.abs().max()) for a, b in zip(la, lb))
# algorithmic flops of the row-cached sampler per sample grid (vertical rows incl. refresh, horizontal columns <= j, head)
print(json.dumps(res))

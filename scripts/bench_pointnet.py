"""PointNet encoder (C = 4, P = 3000, B clouds): the tcgen05 path (v2 pipeline; DVQ_PN_TC_V1=1 selects the round-1 kernel) and
its agreement with the FP32 kernel.  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch, dvq
B, P = int(os.environ.get("B", 4096)), int(os.environ.get("P", 3000))
dev = torch.device("cuda")
torch.manual_seed(0)
enc = dvq.PointNetEncoder(channel=4).to(dev).eval().requires_grad_(False)
g = torch.Generator(device=dev).manual_seed(3000)
obj = 0.1 * torch.randn(B, 4, P, device=dev, generator=g)
obj[:, 3, :] = 0.05 + 0.25 * torch.rand(B, 1, device=dev, generator=g)
for _ in range(2): f, t, _ = enc(obj)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): f, t, _ = enc(obj)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
enc.precision = "fp32"
n = min(B, 256)
f32, t32, _ = enc(obj[:n].contiguous())
print(json.dumps({"tag": os.environ.get("TAG", ""), "B": B, "P": P, "ms": ms, "clouds_per_s": B / ms * 1e3, "tflops": B * P * 558080.0 / ms / 1e9,
                  "max_rel_diff_vs_fp32": float((f[:n] - f32).abs().max() / f32.abs().max()), "trans_diff": float((t[:n] - t32).abs().max())}))

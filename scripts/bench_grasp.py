"""BASELINE config 3 (1 GPU): grasp-generation inference, batch 4096, 3000-point synthetic clouds, random-init
weights: two object PointNets + object-codebook VQ lookup + six part-codebook gathers + decoder MLP + hand stub
+ 778-point PointNet + pose decoder.  Prints one JSON line with grasps/s; writes gpurun_out/grasp.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch, dvq
B, P = int(os.environ.get("B", 4096)), 3000
dev = torch.device("cuda")
torch.manual_seed(0)
net = dvq.GraspGenerator().to(dev).eval()
prec = os.environ.get("PN_PRECISION", "fp16_tc")
for m in (net.obj_encoder_type, net.obj_encoder_pos, net.recon_encoder):
    m.precision = prec
g = torch.Generator(device=dev).manual_seed(3000)
obj = 0.1 * torch.randn(B, 4, P, device=dev, generator=g)
obj[:, 3, :] = (0.05 + 0.25 * torch.rand(B, 1, device=dev, generator=g))
for _ in range(2): net.gen(obj)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters): recon, pos = net.gen(obj)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = B * (2 * 3000 * 558080.0 + 778 * (2 * (3 * 64 + 64 * 128 + 128 * 1024) * 2.0))
line = {"metric": "grasps_per_sec", "value": B / ms * 1e3, "unit": "grasps/s", "ms_per_batch": ms, "batch": B, "points": P,
        "pointnet_tflops": flop / ms / 1e9, "pointnet_precision": prec, "prior": "uniform codes (PixelCNN sampler is a 'next' row)", "hand_layer": "linear stub",
        "finite": bool(torch.isfinite(recon).all() and torch.isfinite(pos).all())}
print(json.dumps(line))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(line, open(os.path.join(ROOT, "gpurun_out", "grasp.json"), "w"), indent=1)

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

# ---- with the PixelCNN prior (gen_net.py:34: GatedPixelCNN(512, 512, 15)), random init, classes limited to the 128 codebook rows
if os.environ.get("PIXELCNN", "1") == "1":
    from dvq.pixelcnn import GatedPixelCNN
    from dvq.grasp import pixelcnn_prior
    import torch.nn.functional as F
    torch.manual_seed(1)
    pcnn = GatedPixelCNN(512, 512, 15).to(dev).eval()
    label = torch.randint(0, 128, (B,), device=dev)
    res = {}

    def timed(fn, iters=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    @torch.no_grad()
    def reference_style_generate():
        # the reference algorithm (models.py:187-197) on the same module: one full forward per sampled position
        x = torch.zeros((B, 3, 3), dtype=torch.int64, device=dev)
        for i in range(3):
            for j in range(3):
                logits = pcnn(x, label)[:, :, i, j].clone()
                logits[:, 128:] = float("-inf")
                probs = F.softmax(logits, -1)
                probs = probs / probs.sum()
                x[:, i, j] = probs.multinomial(1).squeeze(-1)
        return x

    res["sampler_reference_style_ms"] = timed(reference_style_generate)          # cuDNN convs (TF32 by default)
    for prec_p in ("fp32", "tf32"):
        pcnn.precision = prec_p
        res["sampler_row_cached_%s_ms" % prec_p] = timed(lambda: pcnn.generate(None, label, batch_size=B, n_valid=128))
    pcnn.precision = "tf32"
    net.prior = pixelcnn_prior(pcnn, n_valid=128)
    ms2 = timed(lambda: net.gen(obj))
    res.update({"metric": "grasps_per_sec", "value": B / ms2 * 1e3, "unit": "grasps/s", "ms_per_batch": ms2, "batch": B,
                "prior": "GatedPixelCNN(512,512,15) row-cached sampler, tf32 GEMMs, random init, 128 valid classes",
                "pointnet_precision": prec})
    print(json.dumps(res))
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "grasp_pixelcnn.json"), "w"), indent=1)

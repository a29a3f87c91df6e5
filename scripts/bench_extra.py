"""Secondary measurements (not the headline bench line): PointNet encoder throughput, VQ inference path,
codebook sweep of BASELINE config 4 on one GPU, real-model codebook shapes.  Writes gpurun_out/extra.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import numpy as np, torch
import dvq
from dvq import _cabi
from oracle import pointnet_oracle as po, ref_port_torch as port

dev = torch.device("cuda")
out = {}

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

# ---- PointNet encoder ------------------------------------------------------------------------------
sd = po.make_state(5, 4)
net = dvq.PointNetEncoder(channel=4); net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}); net = net.to(dev).eval()
sdt = {k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in sd.items()}
for B in (64, 512, 4096):
    x = torch.from_numpy(po.make_cloud(6, min(B, 64), 4, 3000)).to(dev)
    x = x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1)[:B].contiguous()
    ms = timeit(lambda: net(x), iters=3 if B >= 4096 else 5)
    flop = B * 3000 * 558080.0
    rec = {"ms": ms, "clouds_per_s": B / ms * 1e3, "tflops": flop / ms / 1e9, "hbm_gbs_algorithmic": B * (16 * 3000 + 4096 + 36) / ms / 1e6}
    if B <= 512:   # stock torch ops on the same GPU (the reference's op sequence; cuDNN TF32 convs by default)
        with torch.no_grad():
            ms_t = timeit(lambda: port.pointnet_eval(x, sdt), iters=3)
            ft, tt, _ = port.pointnet_eval(x, sdt)
            fo, to_, _ = net(x)
        rec.update({"torch_ops_ms": ms_t, "speedup_vs_torch_ops": ms_t / ms,
                    "max_abs_diff_vs_torch_ops_tf32": float((fo - ft).abs().max()), "feat_scale": float(ft.abs().max())})
    net.precision = "fp16_tc"
    ms_tc = timeit(lambda: net(x), iters=5)
    with torch.no_grad():
        f_tc, _, _ = net(x)
    net.precision = "fp32"
    with torch.no_grad():
        f_32, _, _ = net(x)
    rec.update({"fp16_tc_ms": ms_tc, "fp16_tc_clouds_per_s": B / ms_tc * 1e3, "fp16_tc_tflops": flop / ms_tc / 1e9,
                "fp16_tc_max_abs_diff_vs_fp32_kernel": float((f_tc - f_32).abs().max()), "feat_scale_fp32": float(f_32.abs().max())})
    out["pointnet_B%d_P3000_C4" % B] = rec
    print("pointnet", B, rec, flush=True)

# ---- VQ: inference path at config 2, real-model shapes, config-4 sweep ---------------------------------
def vq_case(N, K, D, train, path=_cabi.DVQ_PATH_AUTO, iters=5):
    g = torch.Generator(device=dev).manual_seed(4000 + K + D)
    m = dvq.VectorQuantizer(K, D, 0.25, 1.0).to(dev); m.path = path; m.onehot_limit_bytes = 0
    with torch.no_grad():
        m.embedding.weight.copy_((torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K)
        z = torch.randn(N, D, device=dev, generator=g)
        ms = timeit(lambda: m(z, train), iters=iters)
    ref, err = m.last_counters(N)
    return {"N": N, "K": K, "D": D, "train": train, "ms": ms, "latents_per_s": N / ms * 1e3, "tflops_algorithmic": 2.0 * N * K * D / ms / 1e9,
            "hbm_gbs_algorithmic": (N * (8 * D + 8) + 4 * K * D) / ms / 1e6, "rows_refined": ref, "tc_error": err,
            "kernel": "tcgen05+refine" if _cabi.vq_workspace_bytes(N, K, D, 0) > _cabi.vq_workspace_bytes(N, K, D, _cabi.DVQ_PATH_SIMT) else "fp32"}

out["vq_config2_inference"] = vq_case(4194304, 512, 64, False)
out["vq_config2_train"] = vq_case(4194304, 512, 64, True)
out["vq_config2_train_fp32_kernel"] = vq_case(4194304, 512, 64, True, _cabi.DVQ_PATH_SIMT, iters=3)
out["vq_part_codebook_K128_D256_B4096"] = vq_case(4096, 128, 256, False)
out["vq_object_codebook_K128_D1024_B4096"] = vq_case(4096, 128, 1024, False)
for k, v in list(out.items()):
    if k.startswith("vq_"): print(k, v, flush=True)
sweep = []
for K in (512, 1024, 2048, 4096, 8192, 16384):
    for D in (64, 128, 256, 512):
        N = 1 << 20 if K * D <= 512 * 128 else (1 << 18 if K * D <= 4096 * 128 else 1 << 16)
        r = vq_case(N, K, D, True, iters=2)
        sweep.append(r); print("sweep", r, flush=True)
out["vq_config4_sweep_1gpu"] = sweep
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "extra.json"), "w"), indent=1)

"""BASELINE config 4 at full size: codebook sweep K = 512..16384, e_dim = 64..512, N = 16 777 216 latents in total
(N / world per GPU; `python scripts/bench_sweep_full.py` on one GPU, or under torchrun for 2 / 4 / 8), train path
(idx + z_q + loss + perplexity + histogram), default-init codebook (variant A) — the hardest case for the filter.
Per shape: whole-call time (CUDA events, max over ranks), the library's per-stage event times (filter kernel,
exact refine), algorithmic TFLOP/s and GB/s and both roofline fractions.  Writes gpurun_out/sweep_full[_nG].json.

    --rows N_TOTAL   total rows (default 16 777 216)      --shapes K:D,K:D,...   subset of the sweep
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch
import torch.distributed as tdist
import dvq
from dvq import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=16777216)
ap.add_argument("--shapes", default="")
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    tdist.init_process_group("nccl", device_id=dev)
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6553.6)); bf16 = float(peaks.get("bf16_tflops", 1640.2))

shapes = [(K, D) for K in (512, 1024, 2048, 4096, 8192, 16384) for D in (64, 128, 256, 512)]
if args.shapes:
    shapes = [tuple(int(v) for v in s.split(":")) for s in args.shapes.split(",")]
N = args.rows // world
out = []
z = None
zD = 0
for K, D in sorted(shapes, key=lambda s: (s[1], s[0])):
    g = torch.Generator(device=dev).manual_seed(4000 + K + D + 7919 * rank)
    cg = torch.Generator(device=dev).manual_seed(4000 + K + D)
    fresh = zD != D
    if fresh:
        del z
        torch.cuda.empty_cache()
        z = torch.randn(N, D, device=dev, generator=g); zD = D
    m = dvq.VectorQuantizer(K, D, 0.25, 1.0).to(dev); m.onehot_limit_bytes = 0
    with torch.no_grad():
        m.embedding.weight.copy_((torch.rand(K, D, device=dev, generator=cg) * 2 - 1) / K)
    if world > 1:
        dvq.dist.shard_module(m)
    with torch.no_grad():
        for _ in range(3 if fresh else 1):   # the first shape of a row width also fills the caching allocator (and warms NCCL's channels up)
            m(z, True)
        torch.cuda.synchronize(dev)
        if world > 1:
            tdist.barrier(); torch.cuda.synchronize(dev)
        # two timed passes, the faster one is kept: the first pass of a new row width can still carry caching-allocator
        # growth (one 10.8 ms outlier against 6.0 ms for K = 512, e_dim 128 in an earlier run of this script)
        best = None
        for _pass in range(2):
            _cabi.lib.dvq_profile_enable(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                r = m(z, True)
            e1.record(); torch.cuda.synchronize(dev)
            pm, _ = _cabi.profile_mean()
            _cabi.lib.dvq_profile_enable(0)
            t = e0.elapsed_time(e1) / args.iters
            if best is None or t < best[0]:
                best = (t, pm)
        ms, prof_ms = best
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64); tdist.all_reduce(t, op=tdist.ReduceOp.MAX); ms = float(t.item())
    refined, err = m.last_counters(N)
    flop = 2.0 * N * world * K * D
    byts = world * (N * (8 * D + 8) + 4 * K * D)
    rec = {"K": K, "D": D, "N_total": N * world, "n_gpus": world, "ms": ms, "filter_ms": prof_ms[1], "refine_ms": prof_ms[2],
           "latents_per_s": N * world / ms * 1e3, "tflops": flop / ms / 1e9, "gbs": byts / ms / 1e6,
           "tensor_frac_of_bf16_peak_per_gpu": flop / ms / 1e9 / world / bf16, "hbm_frac_per_gpu": byts / ms / 1e6 / world / hbm,
           "filter_tflops_per_gpu": flop / world / prof_ms[1] / 1e9 if prof_ms[1] else None,
           "rows_refined_rank0": refined, "tc_error": err, "loss": float(r[0].item()), "perplexity": float(r[2].item())}
    out.append(rec)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    del m
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"peaks": {"hbm_gbs": hbm, "bf16_tflops": bf16}, "sweep": out},
              open(os.path.join(ROOT, "gpurun_out", "sweep_full%s.json" % ("" if world == 1 else "_n%d" % world)), "w"), indent=1)
if world > 1:
    tdist.destroy_process_group()

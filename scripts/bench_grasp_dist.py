"""BASELINE config 5: diverse-grasp generation throughput, gen_diverse_grasp_obman-style — N_OBJ synthetic object
clouds x N_GRASP random rotations each (rotation recipe: gen_diverse_grasp_FHAB.py:201-207, Rx @ Ry @ Rz with
uniform Euler angles), batch 4096 per step, objects sharded over the ranks (torchrun, one process per GPU; also
runs on one GPU), weights replicated and random-init (seeded), PixelCNN prior (row-cached sampler, TF32 GEMMs,
128 valid classes), tcgen05 PointNets.  The only collective is the all-reduce of the code-usage histograms
(object codebook [128] + six part codebooks [6,128]) at the end.  Timing: CUDA events over all steps of a rank,
max over ranks.  Prints one JSON line on rank 0 and writes gpurun_out/grasp_dist[_nG].json.

    N_OBJ (default 10000)  N_GRASP (default 100)  B (default 4096)
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch
import torch.distributed as tdist
import dvq
from dvq.pixelcnn import GatedPixelCNN
from dvq.grasp import pixelcnn_prior

rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
N_OBJ, N_GRASP, B, P = int(os.environ.get("N_OBJ", 10000)), int(os.environ.get("N_GRASP", 100)), int(os.environ.get("B", 4096)), 3000
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    tdist.init_process_group("nccl", device_id=dev)

torch.manual_seed(0)                                   # replicated weights
net = dvq.GraspGenerator().to(dev).eval()
for m in (net.obj_encoder_type, net.obj_encoder_pos, net.recon_encoder):
    m.precision = os.environ.get("PN_PRECISION", "fp16_tc")
torch.manual_seed(1)
pcnn = GatedPixelCNN(512, 512, 15).to(dev).eval()
pcnn.precision = os.environ.get("PCNN_PRECISION", "fp16_tc")
net.prior = pixelcnn_prior(pcnn, n_valid=128)

# this rank's objects: contiguous shard of the N_OBJ clouds (seeded per object block so any world size sees the same set)
lo, hi = dvq.dist.shard_bounds(N_OBJ, rank, world)
g = torch.Generator(device=dev).manual_seed(5000 + rank)
n_loc = hi - lo
objs = 0.1 * torch.randn(n_loc, 4, P, device=dev, generator=g)
objs[:, 3, :] = 0.05 + 0.25 * torch.rand(n_loc, 1, device=dev, generator=g)
total_loc = n_loc * N_GRASP
steps = math.ceil(total_loc / B)
hist_obj = torch.zeros(128, dtype=torch.int64, device=dev)
hist_part = torch.zeros(6, 128, dtype=torch.int64, device=dev)


def rotations(n):
    a = torch.rand(n, 3, device=dev, generator=g) * (2 * math.pi)
    c, s = torch.cos(a), torch.sin(a)
    one, zero = torch.ones(n, device=dev), torch.zeros(n, device=dev)
    Rx = torch.stack([one, zero, zero, zero, c[:, 0], -s[:, 0], zero, s[:, 0], c[:, 0]], 1).view(n, 3, 3)
    Ry = torch.stack([c[:, 1], zero, s[:, 1], zero, one, zero, -s[:, 1], zero, c[:, 1]], 1).view(n, 3, 3)
    Rz = torch.stack([c[:, 2], -s[:, 2], zero, s[:, 2], c[:, 2], zero, zero, zero, one], 1).view(n, 3, 3)
    return Rx @ Ry @ Rz


@torch.no_grad()
def step(k):
    first = k * B
    n = min(B, total_loc - first)
    gi = torch.arange(first, first + n, device=dev) // N_GRASP           # object of each grasp (object-major order)
    cloud = objs[gi]                                                      # [n,4,P]
    cloud[:, :3, :] = torch.bmm(rotations(n), cloud[:, :3, :])
    recon, pos = net.gen(cloud)
    hist_obj.index_add_(0, net.last["idx6"].view(-1), torch.ones(n, dtype=torch.int64, device=dev))
    codes = net.last["codes"]
    for i in range(6):
        hist_part[i].index_add_(0, codes[:, i].reshape(-1), torch.ones(n, dtype=torch.int64, device=dev))
    return recon, pos


step(0)                                                # warm-up (histograms are reset below)
hist_obj.zero_(); hist_part.zero_()
torch.cuda.synchronize(dev)
if world > 1:
    tdist.barrier(); torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ok = True
for k in range(steps):
    recon, pos = step(k)
if world > 1:                                          # the one collective of the path
    tdist.all_reduce(hist_obj); tdist.all_reduce(hist_part)
e1.record(); torch.cuda.synchronize(dev)
ok = bool(torch.isfinite(recon).all() and torch.isfinite(pos).all())
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64); tdist.all_reduce(t, op=tdist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    total = N_OBJ * N_GRASP
    p = hist_obj.double() / hist_obj.sum().clamp(min=1)
    line = {"metric": "grasps_per_sec", "value": total / ms * 1e3, "unit": "grasps/s", "n_gpus": world, "objects": N_OBJ, "grasps_per_object": N_GRASP,
            "batch": B, "steps_per_rank": steps, "seconds": ms / 1e3, "scaling": "strong (10^6 grasps in total)",
            "object_code_usage_perplexity": float(torch.exp(-(p * torch.log(p + 1e-10)).sum())),
            "histogram_total": int(hist_obj.sum()), "part_histogram_total": int(hist_part.sum()), "finite": ok,
            "prior": "GatedPixelCNN(512,512,15) row-cached sampler, tf32 GEMMs, random init, 128 valid classes",
            "collective": "all-reduce of the usage histograms (int64 [128] + [6,128])"}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", "grasp_dist%s.json" % ("" if world == 1 else "_n%d" % world)), "w"), indent=1)
if world > 1:
    tdist.destroy_process_group()

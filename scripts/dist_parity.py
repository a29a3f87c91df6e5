"""torchrun -N: the row-sharded VQ path on N GPUs equals the 1-GPU result — indices and z_q per shard
bit-identical, histogram exactly equal, loss / perplexity within 1e-6 (fp64 summation order)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import torch, torch.distributed as tdist
import dvq

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
tdist.init_process_group("nccl", device_id=dev)
N, K, D = 1000003, 512, 64
g = torch.Generator(device=dev).manual_seed(77)            # same seed on every rank: identical full tensors
E = (torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K
z = torch.randn(N, D, device=dev, generator=g)
ok = True
for path in (dvq._cabi.DVQ_PATH_SIMT, dvq._cabi.DVQ_PATH_AUTO):
    vq1 = dvq.VectorQuantizer(K, D, 0.25, 1.0).to(dev); vq1.path = path; vq1.onehot_limit_bytes = 0
    vqd = dvq.VectorQuantizer(K, D, 0.25, 1.0).to(dev); vqd.path = path; vqd.onehot_limit_bytes = 0
    with torch.no_grad():
        vq1.embedding.weight.copy_(E); vqd.embedding.weight.copy_(E)
        l1, q1, p1, _, i1 = vq1(z, True)                      # whole tensor on this GPU
        dvq.dist.shard_module(vqd)
        lo, hi = dvq.dist.shard_bounds(N, rank, world)
        ld, qd, pd, _, idd = vqd(z[lo:hi].contiguous(), True)
    same_idx = bool(torch.equal(idd, i1[lo:hi])); same_q = bool(torch.equal(qd, q1[lo:hi]))
    same_hist = bool(torch.equal(vqd.last_stats[:K], vq1.last_stats[:K]))
    rl = abs(ld.item() - l1.item()) / abs(l1.item()); rp = abs(pd.item() - p1.item()) / abs(p1.item())
    good = same_idx and same_q and same_hist and rl < 1e-6 and rp < 1e-6
    ok = ok and good
    if rank == 0:
        print("path", path, "world", world, "idx", same_idx, "zq", same_q, "hist", same_hist, "loss_rel", rl, "ppl_rel", rp, "OK" if good else "FAIL")
t = torch.tensor([1 if ok else 0], device=dev); tdist.all_reduce(t, op=tdist.ReduceOp.MIN)
if rank == 0:
    print("DIST_PARITY", "PASS" if int(t.item()) == 1 else "FAIL")
tdist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)

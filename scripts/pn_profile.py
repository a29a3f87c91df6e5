"""Minimal driver for ncu captures of the tcgen05 PointNet trunk: a few forwards at B=512, P=3000."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))
import numpy as np, torch, dvq
from oracle import pointnet_oracle as po
net = dvq.PointNetEncoder(channel=4); net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in po.make_state(5, 4).items()})
net = net.cuda().eval(); net.precision = os.environ.get("PN_PRECISION", "fp16_tc")
x = torch.from_numpy(po.make_cloud(6, 64, 4, 3000)).cuda().repeat(8, 1, 1).contiguous()
for _ in range(3): f, t, _ = net(x)
torch.cuda.synchronize(); print(float(f.abs().max()))

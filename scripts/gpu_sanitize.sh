#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (memcheck; racecheck on the shared-memory pipelines)
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys; sys.path.insert(0, "d-vqvae_b200"); sys.path.insert(0, ".")
import numpy as np, torch, dvq
from dvq import _cabi
torch.manual_seed(0)
for K, D, N in ((512, 64, 1000), (512, 128, 700), (1024, 256, 300), (2048, 512, 260), (2048, 128, 300), (4096, 256, 513)):
    vq = dvq.VectorQuantizer(K, D, 0.25, 1.0).cuda(); vq.onehot_limit_bytes = 0
    z = torch.randn(N, D, device="cuda")
    with torch.no_grad():
        a = vq(z, True); b = vq(z, True); c = vq(z, False)
    torch.cuda.synchronize()
    assert vq.last_counters(N)[1] == 0 and torch.equal(a[4], b[4])
enc = dvq.PointNetEncoder(channel=4).cuda().eval().requires_grad_(False)
f, t, _ = enc(0.1 * torch.randn(5, 4, 700, device="cuda"))
enc3 = dvq.PointNetEncoder(channel=3).cuda().eval().requires_grad_(False)
f3, t3, _ = enc3(0.1 * torch.randn(3, 3, 130, device="cuda"))
rs = np.random.RandomState(0)
w = rs.rand(778, 16) ** 4; jr = rs.rand(16, 778) ** 8
m = {"v_template": 0.1 * rs.randn(778, 3), "shapedirs": 0.01 * rs.randn(778, 3, 10), "posedirs": 0.002 * rs.randn(778, 3, 135),
     "J_regressor": jr / jr.sum(1, keepdims=True), "weights": w / w.sum(1, keepdims=True), "hands_components": rs.randn(45, 45) / 6.7,
     "hands_mean": 0.3 * rs.randn(45), "parents": np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14])}
layer = dvq.ManoLayer(m).cuda()
v = layer(betas=torch.randn(7, 10, device="cuda"), hand_pose=torch.randn(7, 45, device="cuda")).vertices
torch.cuda.synchronize()
print("ok", float(f.abs().max()), float(f3.abs().max()), float(v.abs().max()))
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_small.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done

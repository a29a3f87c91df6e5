"""GPU parity tests of the fused PointNet encoder against the real reference's golden outputs
(CPU FP32) and the numpy oracle.  Tolerance: 5e-5 of the feature scale — the kernel folds BN into
the conv weights and accumulates with sequential-k FP32 FMA, the reference runs conv / BN / ReLU
as separate MKL-DNN ops; both are FP32 with different rounding points (observed ~1e-6)."""
import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as po
from _cases import PN_CASES, load_gold

pytestmark = pytest.mark.gpu
TOL = 5e-5


def _encoder(seed, c, precision="fp32"):
    """The FP32 CUDA-core kernel unless a test asks for the default (tcgen05, "fp16_tc")."""
    import dvq
    net = dvq.PointNetEncoder(global_feat=True, feature_transform=False, channel=c)
    assert net.precision == "fp16_tc"                        # the drop-in default is the tensor-core path
    net.precision = precision
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in po.make_state(seed, c).items()}, strict=True)
    return net.cuda().eval().requires_grad_(False)


@pytest.mark.parametrize("name", PN_CASES)
def test_pointnet_matches_reference_golden(name):
    g = load_gold(name)
    b, c, p, seed = [int(v) for v in g["meta"]]
    net = _encoder(seed, c)
    x = torch.from_numpy(po.make_cloud(seed + 1, b, c, p)).cuda()
    feat, trans, tf = net(x)
    assert tf is None and tuple(feat.shape) == (b, 1024) and tuple(trans.shape) == (b, 3, 3)
    scale = float(np.abs(g["feat"]).max())
    assert np.abs(feat.cpu().numpy() - g["feat"]).max() <= TOL * scale
    assert np.abs(trans.cpu().numpy() - g["trans"]).max() <= TOL * max(1.0, float(np.abs(g["trans"]).max()))


def test_pointnet_batch_independence_and_permutation_invariance():
    """max-pool over the point set: permuting points leaves the feature unchanged bit-for-bit
    only up to the tile-local max order, which is exact for max; clouds are independent."""
    net = _encoder(11, 4)
    x = torch.from_numpy(po.make_cloud(12, 6, 4, 1000)).cuda()
    feat, trans, _ = net(x)
    perm = torch.randperm(1000, device="cuda")
    feat_p, trans_p, _ = net(x[:, :, perm].contiguous())
    assert torch.allclose(trans, trans_p, rtol=0, atol=2e-6)
    assert torch.allclose(feat, feat_p, rtol=0, atol=2e-5)
    f1, t1, _ = net(x[2:3].contiguous())
    assert torch.equal(f1, feat[2:3]) and torch.equal(t1, trans[2:3])
    # non-contiguous input (the permute at gen_net.py:120)
    verts = x[:, :3, :].permute(0, 2, 1).contiguous()       # [B,P,3]
    net3 = _encoder(13, 3)
    fa, _, _ = net3(verts.permute(0, 2, 1))
    fb, _, _ = net3(verts.permute(0, 2, 1).contiguous())
    assert torch.equal(fa, fb)


def test_pointnet_vs_oracle_larger_batch():
    net = _encoder(21, 4)
    xn = po.make_cloud(22, 9, 4, 3000)
    feat, trans, _ = net(torch.from_numpy(xn).cuda())
    rf, rt, _ = po.pointnet_forward(xn, po.make_state(21, 4))
    assert np.abs(feat.cpu().numpy() - rf).max() <= TOL * float(np.abs(rf).max())
    assert np.abs(trans.cpu().numpy() - rt).max() <= TOL * max(1.0, float(np.abs(rt).max()))


def test_pointnet_refuses_training_mode_and_bad_shapes():
    import dvq
    net = dvq.PointNetEncoder(channel=4).cuda().requires_grad_(False)
    with pytest.raises(RuntimeError, match="eval"):
        net(torch.randn(1, 4, 8, device="cuda"))
    net.eval()
    with pytest.raises(ValueError):
        net(torch.randn(1, 3, 8, device="cuda"))
    f, t, _ = net(torch.randn(2, 4, 1, device="cuda"))       # a single point
    assert tuple(f.shape) == (2, 1024)
    # the fold cache notices in-place weight updates
    f0, _, _ = net(torch.ones(1, 4, 5, device="cuda"))
    with torch.no_grad():
        net.conv3.weight.mul_(2.0)
    f1, _, _ = net(torch.ones(1, 4, 5, device="cuda"))
    assert not torch.equal(f0, f1)
    # an edit through .data does not bump the version counter: invalidate() is the documented way
    net.conv3.weight.data.mul_(0.5)
    net.invalidate()
    f2, _, _ = net(torch.ones(1, 4, 5, device="cuda"))
    assert torch.allclose(f2, f0, rtol=1e-3, atol=1e-5)
    with pytest.raises(RuntimeError, match="inference-only"):
        net(torch.ones(1, 4, 5, device="cuda", requires_grad=True))


TOL_TC = 1e-3   # FP16 operands (10-bit mantissa, = the TF32 cuDNN path of the reference on a GPU), FP32 accumulation; measured 2.6e-4


@pytest.mark.parametrize("name", PN_CASES)
def test_pointnet_fp16_tc_matches_reference_golden(name):
    g = load_gold(name)
    b, c, p, seed = [int(v) for v in g["meta"]]
    net = _encoder(seed, c)
    net.precision = "fp16_tc"
    x = torch.from_numpy(po.make_cloud(seed + 1, b, c, p)).cuda()
    feat, trans, tf = net(x)
    assert tf is None and tuple(feat.shape) == (b, 1024)
    scale = float(np.abs(g["feat"]).max())
    assert np.abs(feat.cpu().numpy() - g["feat"]).max() <= TOL_TC * scale
    assert np.abs(trans.cpu().numpy() - g["trans"]).max() <= TOL_TC * max(1.0, float(np.abs(g["trans"]).max()))


def test_pointnet_fp16_tc_vs_fp32_kernel_many_clouds():
    """More clouds than CTAs-groups, ragged point count: the persistent cloud loop and the tail tile."""
    net = _encoder(31, 4)
    x = torch.from_numpy(po.make_cloud(32, 70, 4, 1037)).cuda()
    f32, t32, _ = net(x)
    net.precision = "fp16_tc"
    f16, t16, _ = net(x)
    scale = float(f32.abs().max())
    assert float((f16 - f32).abs().max()) <= TOL_TC * scale
    assert float((t16 - t32).abs().max()) <= TOL_TC * max(1.0, float(t32.abs().max()))
    f16b, _, _ = net(x[5:6].contiguous())
    assert torch.equal(f16b, f16[5:6])                     # clouds are independent, result is deterministic


@pytest.mark.parametrize("scale,c", [(40.0, 4), (1e-3, 3), (1.0, 3)])
def test_pointnet_fp16_tc_layer1_keeps_fp32_accuracy_over_coordinate_ranges(scale, c):
    """Layer 1 runs on the tensor pipe with the coordinates and weights as hi/lo FP16 pairs (x.w + b to ~2^-22): large and tiny
    coordinates (FP16 alone would lose them to its 11-bit mantissa / subnormal range) still agree with the FP32 kernel to the
    same bar, for 3- and 4-channel clouds and more than one tile stream per CTA group."""
    net = _encoder(33, c)
    x = torch.from_numpy(po.make_cloud(34, 150, c, 700)).cuda() * scale
    f32, t32, _ = net(x)
    net.precision = "fp16_tc"
    f16, t16, _ = net(x)
    fs = float(f32.abs().max())
    assert float((f16 - f32).abs().max()) <= TOL_TC * fs
    assert float((t16 - t32).abs().max()) <= TOL_TC * max(1.0, float(t32.abs().max()))

"""GPU: the batched grasp-generation graph (GenNet.gen restated around the B200 modules) against a
stage-by-stage recomputation with the oracle on the same weights, codes and hand stub."""
import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as po
from oracle import vq_oracle as vo

pytestmark = pytest.mark.gpu


def test_grasp_pipeline_matches_oracle_stages():
    import dvq
    torch.manual_seed(0)
    net = dvq.GraspGenerator().cuda().eval()
    for m in (net.obj_encoder_type, net.obj_encoder_pos, net.recon_encoder):
        m.precision = "fp32"                                  # stage-by-stage FP32 oracle bars below
    for name in ("obj_encoder_type", "obj_encoder_pos"):
        getattr(net, name).load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in po.make_state(hash(name) % 1000, 4).items()})
    net.recon_encoder.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in po.make_state(77, 3).items()})
    x = po.make_cloud(9, 5, 4, 777)
    recon, pos = net.gen(torch.from_numpy(x).cuda())
    assert tuple(recon.shape) == (5, 55) and tuple(pos.shape) == (5, 6)
    L = net.last
    sd_t = {k: v.detach().cpu().numpy() for k, v in net.obj_encoder_type.state_dict().items()}
    rf, _, _ = po.pointnet_forward(x, sd_t)
    assert np.abs(L["feat_type"].cpu().numpy() - rf).max() <= 5e-5 * np.abs(rf).max()
    E6 = net.vqvae6.vector_quantization.embedding.weight.detach().cpu().numpy()
    ridx, rzq = vo.forward_infer(L["feat_type"].cpu().numpy(), E6)
    assert vo.allowed_index_mismatch(L["feat_type"].cpu().numpy(), E6, L["idx6"].cpu().numpy(), ridx)[1] == 0
    assert np.array_equal(L["obj_emb"].cpu().numpy().view(np.uint32), E6[L["idx6"].cpu().numpy().reshape(-1)].view(np.uint32))
    codes = L["codes"].cpu().numpy()
    embs = [getattr(net, "vqvae%d" % i).vector_quantization.embedding.weight.detach().cpu().numpy()[codes[:, i]] for i in range(6)]
    z_out = torch.from_numpy(np.concatenate(embs + [L["feat_type"].cpu().numpy()], axis=1))
    ref_recon = net.decoder.cpu()(z_out)
    assert torch.allclose(recon.cpu(), ref_recon, rtol=1e-4, atol=1e-5)
    sd_r = {k: v.detach().cpu().numpy() for k, v in net.recon_encoder.state_dict().items()}
    hf, _, _ = po.pointnet_forward(np.ascontiguousarray(L["verts"].permute(0, 2, 1).cpu().numpy()), sd_r)
    assert np.abs(L["hand_feat"].cpu().numpy() - hf).max() <= 5e-5 * np.abs(hf).max()


def test_grasp_state_dict_names_follow_reference():
    import dvq
    keys = set(dvq.GraspGenerator().state_dict())
    for k in ("obj_encoder_type.stn.conv1.weight", "vqvae0.vector_quantization.embedding.weight", "vqvae6.vector_quantization.embedding.weight",
              "decoder.MLP.L0.weight", "decoder.MLP.L2.bias", "recon_encoder.bn3.running_var", "pos_decoder.MLP.L1.weight"):
        assert k in keys, k


def test_grouped_gather_and_graphed_generation():
    """dvq_gather_multi == six get_embbeding calls; gen_graphed (one CUDA-graph launch per batch) == gen for the
    deterministic stages, with the kernel-backed PixelCNN prior."""
    import dvq
    from dvq.grasp import pixelcnn_prior
    from dvq.pixelcnn import GatedPixelCNN
    torch.manual_seed(0)
    pcnn = GatedPixelCNN(512, 256, 2, 128).cuda().eval().requires_grad_(False)
    pcnn.precision = "fp16_tc"
    net = dvq.GraspGenerator(prior=pixelcnn_prior(pcnn, n_valid=128)).cuda().eval().requires_grad_(False)
    obj = 0.1 * torch.randn(130, 4, 500, device="cuda")
    recon, pos = net.gen(obj)
    L = net.last
    embs = [getattr(net, "vqvae%d" % i).get_embbeding(L["codes"][:, i].contiguous(), 256) for i in range(6)]
    z_ref = torch.cat(embs + [L["feat_type"]], dim=1)
    assert torch.allclose(net.decoder(z_ref).view(130, 55), recon, rtol=1e-5, atol=1e-6)
    r2, p2 = net.gen_graphed(obj)
    L2 = net.last
    assert torch.equal(L2["feat_type"], L["feat_type"]) and torch.equal(L2["idx6"], L["idx6"])     # stages before the sampler
    assert r2.shape == recon.shape and torch.isfinite(r2).all() and torch.isfinite(p2).all()
    r3, _ = net.gen_graphed(obj * 1.01)                    # replay with new input contents
    assert torch.isfinite(r3).all()
    pcnn._tc_sampler.check()


def _fixed_prior(codes):
    def prior(idx6, batch):
        g = torch.from_numpy(codes).to(idx6.device)
        return torch.stack([g[:, 0, 1], g[:, 0, 2], g[:, 1, 1], g[:, 1, 2], g[:, 2, 1], g[:, 2, 2]], dim=1)
    return prior


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("fp16_tc", 5e-3)])
def test_generation_graph_matches_the_real_reference_gen(precision, tol):
    """tests/golden/gennet_stages.npz holds the outputs of the REAL reference's GenNet.gen (network/gen_net.py:78-125,
    run on CPU by oracle/gen_golden_gennet.py with these weights, fixed prior codes and the linear hand stub), one
    object at a time (reference semantics B = 1).  The batched B200 graph must reproduce them: object-codebook index
    exactly, the 55 hand parameters and the 6-DoF pose within `tol` of their scale."""
    import hashlib
    import dvq
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "gennet_stages.npz"))
    torch.manual_seed(0)
    net = dvq.GraspGenerator()
    h = hashlib.sha256()
    for k, v in sorted(net.state_dict().items()):
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    if h.hexdigest() != str(g["sd_sha256"]):
        pytest.skip("torch CPU RNG stream differs from the build container")
    net.prior = _fixed_prior(g["codes"])
    net = net.cuda().eval().requires_grad_(False)
    for m in (net.obj_encoder_type, net.obj_encoder_pos, net.recon_encoder):
        m.precision = precision
    obj = torch.from_numpy(po.make_cloud(int(g["cloud_seed"]), 3, 4, 3000)).cuda()
    recon, pos = net.gen(obj)                                     # all three objects in ONE batch
    L = net.last
    fscale = float(np.abs(g["feat_type"]).max())
    assert np.abs(L["feat_type"].cpu().numpy() - g["feat_type"]).max() <= tol * fscale
    assert np.array_equal(L["idx6"].cpu().numpy().reshape(-1), g["idx6"])
    assert np.abs(recon.cpu().numpy() - g["recon"]).max() <= tol * max(1.0, float(np.abs(g["recon"]).max()))
    assert np.abs(pos.cpu().numpy() - g["recon_pos"]).max() <= tol * max(1.0, float(np.abs(g["recon_pos"]).max()))

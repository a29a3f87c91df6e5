"""CPU: pins oracle/ against the golden vectors produced by the real reference
(oracle/gen_golden.py).  Index rule: SURVEY §8c near-tie band (FP64 gap < 1e-6 rel);
z_q bit-exact given the index; loss/perplexity within 1e-5 relative."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import pointnet_oracle as po
from oracle import vq_oracle as vo
from _cases import GOLD, PN_CASES, VQ_CASES, load_gold, rel_err, vq_inputs

REL = 1e-5


@pytest.mark.parametrize("name", list(VQ_CASES))
def test_vq_oracle_matches_reference_outputs(name):
    z, E, al, beta = vq_inputs(name)
    g = load_gold(name)
    loss, zq, ppl, enc, idx = vo.forward_train(z, E, al, beta)
    idx_i, zq_i = vo.forward_infer(z, E)
    assert idx.shape == g["idx_train"].shape and idx.dtype == np.int64
    assert zq.shape == z.shape == g["zq_train"].shape
    n_mis, n_bad, worst = vo.allowed_index_mismatch(z, E, idx, g["idx_train"])
    assert n_bad == 0, (n_mis, worst)
    n_mis_i, n_bad_i, _ = vo.allowed_index_mismatch(z, E, idx_i, g["idx_infer"])
    assert n_bad_i == 0
    # z_q bit-exact given the index: recompute from the REFERENCE's index
    ridx = g["idx_train"].astype(np.int64)
    assert np.array_equal(vo.zq_infer_from_idx(E, ridx).reshape(z.shape).view(np.uint32), g["zq_infer"].view(np.uint32))
    assert np.array_equal(vo.zq_train_from_idx(z, E, ridx).reshape(z.shape).view(np.uint32), g["zq_train"].view(np.uint32))
    # rows where the index agrees must agree bitwise in z_q too
    same = (idx.reshape(-1) == ridx.reshape(-1))
    assert np.array_equal(zq.reshape(-1, E.shape[1])[same].view(np.uint32),
                          g["zq_train"].reshape(-1, E.shape[1])[same].view(np.uint32))
    assert rel_err(loss, g["loss"]) < REL
    assert rel_err(ppl, g["perplexity"]) < REL or n_mis > 0 and rel_err(ppl, g["perplexity"]) < 1e-3
    assert enc.shape == (idx.shape[0], E.shape[0]) and enc.dtype == np.float32
    assert np.array_equal(enc.sum(0).astype(np.int64), np.bincount(idx.reshape(-1), minlength=E.shape[0]))
    # stats form used by the multi-GPU path
    cidx, hist, sse = vo.forward_stats_chunked(z, E, chunk=100)
    assert np.array_equal(cidx, idx.reshape(-1))
    assert rel_err(vo.loss_from_sse(sse, idx.shape[0], E.shape[1], al, beta), g["loss"]) < REL
    if n_mis == 0:
        assert np.array_equal(hist, g["hist"])
        assert rel_err(vo.perplexity_from_hist(hist, idx.shape[0]), g["perplexity"]) < REL


def test_exact_ties_pick_lowest_index():
    z, E, al, beta = vq_inputs("vq_dupes")
    g = load_gold("vq_dupes")
    assert g["idx_train"].max() < E.shape[0] // 2          # the reference itself: duplicates never win
    idx, _ = vo.forward_infer(z, E)
    assert idx.max() < E.shape[0] // 2


def test_get_emb_and_wrapper_golden():
    z, E, al, beta = vq_inputs("vq_k128_d256")
    g = load_gold("vq_get_emb")
    assert np.array_equal(vo.get_emb(E, g["picks"]).view(np.uint32), g["embs"].view(np.uint32))
    loss, _, ppl, _, idx = vo.forward_train(z, E, al, beta)
    assert rel_err(loss, g["wrapper_loss"]) < REL and rel_err(ppl, g["wrapper_ppl"]) < REL
    assert vo.allowed_index_mismatch(z, E, idx, g["wrapper_idx"])[1] == 0


def test_config1_known_answer():
    """BASELINE.md §2 anchor: loss 1.248528003692627, perplexity 455.29669189453125."""
    import torch
    g = load_gold("vq_config1_full")
    torch.manual_seed(0)
    emb = torch.nn.Embedding(512, 64)
    emb.weight.data.uniform_(-1.0 / 512, 1.0 / 512)
    z = torch.randn(65536, 64).numpy()
    E = emb.weight.detach().numpy()
    if hashlib.sha256(z.tobytes()).hexdigest() != str(g["z_sha256"]):
        pytest.skip("torch CPU RNG stream differs from the build container")
    assert hashlib.sha256(E.tobytes()).hexdigest() == str(g["codebook_sha256"])
    assert abs(float(g["loss"]) - 1.248528003692627) < 1e-9 and abs(float(g["perplexity"]) - 455.29669189453125) < 1e-6
    idx, hist, sse = vo.forward_stats_chunked(z, E)
    n_mis, n_bad, worst = vo.allowed_index_mismatch(z, E, idx, g["idx"].astype(np.int64))
    assert n_bad == 0 and n_mis < 200, (n_mis, worst)
    assert rel_err(vo.loss_from_sse(sse, 65536, 64, 1, 0.25), g["loss"]) < REL
    assert rel_err(vo.perplexity_from_hist(hist, 65536), g["perplexity"]) < REL
    zq = vo.zq_train_from_idx(z, E, g["idx"].astype(np.int64))
    assert hashlib.sha256(zq.tobytes()).hexdigest() == str(g["zq_train_sha256"])


@pytest.mark.parametrize("name", PN_CASES)
def test_pointnet_oracle_matches_reference_outputs(name):
    g = load_gold(name)
    b, c, p, seed = [int(v) for v in g["meta"]]
    sd = po.make_state(seed, c)
    x = po.make_cloud(seed + 1, b, c, p)
    feat, trans, tf = po.pointnet_forward(x, sd)
    assert tf is None and feat.shape == (b, 1024) and trans.shape == (b, 3, 3)
    scale = float(np.abs(g["feat"]).max())
    assert np.abs(feat - g["feat"]).max() <= 2e-5 * scale
    assert np.abs(trans - g["trans"]).max() <= 2e-5 * max(1.0, float(np.abs(g["trans"]).max()))


def test_torch_port_is_bit_identical_to_reference():
    with open(os.path.join(GOLD, "port_vs_reference.json")) as f:
        rep = json.load(f)
    for name, r in rep.items():
        if name.startswith("vq_"):
            assert all(r.values()), (name, r)
        else:
            assert r["feat_max_abs_diff"] == 0.0 and r["trans_max_abs_diff"] == 0.0

"""CPU: the C-ABI library loads and exports every symbol include/dvq.h declares; argument
validation that needs no device works; the product path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dvq.h")).read()
    return sorted(set(re.findall(r"DVQ_API\s+[\w\s\*]+?\b(dvq_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dvq import _cabi
    names = _declared()
    assert len(names) >= 14
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_cabi.SYMBOLS), (set(names) ^ set(_cabi.SYMBOLS))
    assert lib.dvq_abi_version() == _cabi.ABI_VERSION


def test_workspace_queries_and_shape_errors_need_no_device():
    from dvq import _cabi
    assert _cabi.vq_workspace_bytes(4096, 512, 64, 0) >= 512 * 4
    assert _cabi.vq_workspace_bytes(0, 1, 1, _cabi.DVQ_PATH_SIMT) > 0
    assert _cabi.pointnet_workspace_bytes(8, 4, 3000) >= 8 * 1024 * 4 * 2
    with pytest.raises(RuntimeError, match="BAD_SHAPE"):
        _cabi.vq_workspace_bytes(10, 0, 64, 0)
    with pytest.raises(RuntimeError, match="BAD_SHAPE"):
        _cabi.pointnet_workspace_bytes(1, 5, 100)
    sz = ctypes.c_size_t()
    assert _cabi.lib.dvq_vq_workspace_bytes(1 << 40, 8, 8, 0, ctypes.byref(sz)) == -1
    assert "2^31" in _cabi.last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_cpu_fallback():
    import dvq
    from dvq import _cabi
    vq = dvq.VectorQuantizer(16, 8, 0.25, 1)
    with pytest.raises(ValueError, match="no CPU path"):
        vq(torch.randn(4, 8), True)
    with pytest.raises(ValueError, match="no CPU path"):
        vq(torch.randn(4, 8), False)
    with pytest.raises(ValueError, match="no CPU path"):
        vq.get_emb(torch.tensor([1]), 8)
    pn = dvq.PointNetEncoder(channel=4).eval()
    with pytest.raises(ValueError, match="no CPU path"):
        pn(torch.randn(1, 4, 16))
    # straight through the C ABI: the library refuses, it does not compute on the host
    buf = (ctypes.c_float * 64)()
    idx = (ctypes.c_int64 * 4)()
    ws = ctypes.create_string_buffer(4096 + 256)
    wsp = (ctypes.addressof(ws) + 255) // 256 * 256
    rc = _cabi.lib.dvq_vq_forward(ctypes.addressof(buf), ctypes.addressof(buf), 4, 2, 8, 0, ctypes.addressof(buf),
                                  ctypes.addressof(idx), None, None, None, wsp, 4096, None)
    assert rc in (-3, -5), rc


def test_module_surface_matches_reference():
    """Constructor args, attribute names, state_dict keys."""
    import dvq
    vq = dvq.VectorQuantizer(128, 256, 0.25, 1)
    assert (vq.n_e, vq.e_dim, vq.beta, vq.al) == (128, 256, 0.25, 1)
    assert list(vq.state_dict()) == ["embedding.weight"] and vq.embedding.weight.shape == (128, 256)
    assert float(vq.embedding.weight.abs().max()) <= 1.0 / 128
    w = dvq.VQVAE(128, 32, 2, 128, 1024, 2, a=0)
    assert list(w.state_dict()) == ["vector_quantization.embedding.weight"]
    assert w.vector_quantization.al == 0 and w.vector_quantization.beta == 2
    pn = dvq.PointNetEncoder(global_feat=True, feature_transform=False, channel=4)
    keys = list(pn.state_dict())
    expected = []
    for pre, fcs in (("stn.", True), ("", False)):
        for c in ("conv1", "conv2", "conv3") + (("fc1", "fc2", "fc3") if fcs else ()):
            expected += [pre + c + ".weight", pre + c + ".bias"]
        for b in ("bn1", "bn2", "bn3") + (("bn4", "bn5") if fcs else ()):
            expected += [pre + b + s for s in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")]
    assert sorted(keys) == sorted(expected)
    assert pn.stn.conv1.weight.shape == (64, 4, 1) and pn.conv3.weight.shape == (1024, 128, 1)
    from oracle import pointnet_oracle as po
    sd = {k: torch.from_numpy(v) for k, v in po.make_state(1, 4).items()}
    pn.load_state_dict(sd, strict=True)          # reference-keyed checkpoint loads unchanged
    with pytest.raises(NotImplementedError):
        dvq.PointNetEncoder(global_feat=False)


def test_bn_fold_matches_oracle_math():
    import numpy as np
    import dvq
    from dvq.pointnet import _fold
    from oracle import pointnet_oracle as po
    sd = po.make_state(3, 4)
    pn = dvq.PointNetEncoder(channel=4).eval()
    pn.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    w, b = _fold(pn.conv2.weight, pn.conv2.bias, pn.bn2)
    x = np.random.RandomState(0).standard_normal((1, 64, 7)).astype(np.float32)
    ref = po._bn_eval(po._conv1x1(x, sd["conv2.weight"], sd["conv2.bias"]), sd, "bn2")
    got = np.einsum("oc,bcp->bop", w.numpy(), x) + b.numpy()[None, :, None]
    assert np.abs(got - ref).max() < 1e-5


def test_tc_shared_memory_plan_per_shape():
    """Host-only view of the tcgen05 VQ kernel's shared-memory plan (dvq_debug_tc_layout): every shape of the
    BASELINE sweep fits the 227 KB budget, and the per-shape choices that were measured on the GPU are pinned —
    resident image with a shared-memory histogram at config 2; a third ring slot instead of the second z staging slot
    from 8 chunks per tile on (e_dim 64); at e_dim 128 the third slot replaces the second A image only from 32
    chunks per tile on (with a single A image K = 2048 / 4096 measured 10-17 % slower)."""
    from dvq import _cabi

    def plan(K, D):
        out = (ctypes.c_int * 8)()
        _cabi.check(_cabi.lib.dvq_debug_tc_layout(K, D, out), "dvq_debug_tc_layout")
        return dict(zip(("ok", "ds", "ns", "a_bufs", "nslots", "nstage", "bytes", "hist"), list(out)))

    for K in (512, 1024, 2048, 4096, 8192, 16384):
        for D in (16, 32, 64, 128, 256, 512):
            p = plan(K, D)
            assert p["ok"] == 1 and p["bytes"] <= 227 * 1024, (K, D, p)
            assert p["ds"] * p["ns"] == D and 2 <= p["nslots"] <= 4 and 1 <= p["nstage"] <= 2 and 1 <= p["a_bufs"] <= 2
    assert plan(512, 64) == dict(ok=1, ds=64, ns=1, a_bufs=2, nslots=2, nstage=2, bytes=plan(512, 64)["bytes"], hist=1)
    assert (plan(1024, 64)["nslots"], plan(1024, 64)["nstage"], plan(1024, 64)["a_bufs"]) == (2, 2, 2)
    for K in (2048, 4096, 16384):
        assert (plan(K, 64)["nslots"], plan(K, 64)["nstage"], plan(K, 64)["a_bufs"]) == (3, 1, 2)
    assert (plan(4096, 128)["nslots"], plan(4096, 128)["a_bufs"]) == (2, 2)
    # few blocks per tile at e_dim 128 / 256: measured layouts (a short tile is bound by its load -> convert -> MMA -> filter chain)
    assert (plan(512, 128)["a_bufs"], plan(512, 128)["nstage"], plan(512, 128)["nslots"]) == (1, 1, 2)   # 17 % ahead of (2, 2, 2), 3 % of (1, 1, 3)
    assert (plan(512, 256)["a_bufs"], plan(512, 256)["nstage"], plan(512, 256)["nslots"]) == (1, 1, 2)   # 14 % ahead of (1, 2, 2)
    assert (plan(1024, 128)["a_bufs"], plan(1024, 128)["nstage"], plan(1024, 128)["nslots"]) == (2, 1, 2)  # 6 % ahead of (2, 2, 2)
    assert (plan(128, 256)["a_bufs"], plan(128, 256)["nstage"], plan(128, 256)["nslots"]) == (1, 2, 2)   # the grasp codebooks: general rule
    assert (plan(16384, 128)["nslots"], plan(16384, 128)["a_bufs"], plan(16384, 128)["nstage"]) == (3, 1, 1)
    assert plan(16384, 256)["a_bufs"] == 1 and plan(16384, 512)["ds"] == 32
    # shapes the tensor-core path does not take
    assert plan(500, 64)["ok"] == 0 and plan(512, 48)["ok"] == 0 and plan(32768, 64)["ok"] == 0


def test_tc_pair_kernel_selection_rule():
    """Host-only view of the CTA-pair (cta_group::2) kernel's selection and shared-memory plan (dvq_debug_tc_pair_layout):
    streamed codebooks at e_dim 512, and at e_dim 128 / 256 from K = 2048 on (the shapes where it measured ahead of
    the single-CTA kernel); never for a resident operand image (config 2) or a single row tile; every selected
    shape's plan fits the 227 KB budget with half-block ring slots (128 codes)."""
    import os
    from dvq import _cabi
    if os.environ.get("DVQ_TC_PAIR"):
        pytest.skip("DVQ_TC_PAIR overrides the rule")

    def plan(N, K, D):
        out = (ctypes.c_int * 8)()
        _cabi.check(_cabi.lib.dvq_debug_tc_pair_layout(N, K, D, out), "dvq_debug_tc_pair_layout")
        return dict(zip(("pair", "ds", "ns", "a_bufs", "nslots", "nstage", "bytes", "bcodes"), list(out)))

    N = 1 << 20
    for K in (512, 1024, 2048, 4096, 8192, 16384):
        for D in (64, 128, 256, 512):
            p = plan(N, K, D)
            want = D == 512 or (D >= 128 and K >= 2048)
            assert p["pair"] == int(want), (K, D, p)
            if want:
                assert p["bytes"] <= 227 * 1024 and p["bcodes"] == 128 and 2 <= p["nslots"] <= 8 and p["ds"] * p["ns"] == D
    assert plan(N, 512, 64)["pair"] == 0              # config 2: resident image
    assert plan(128, 16384, 512)["pair"] == 0         # one row tile: nothing to pair
    assert plan(129, 16384, 512)["pair"] == 1
    assert plan(N, 128, 256)["pair"] == 0             # the grasp codebooks (K = 128): single-CTA kernel


@pytest.mark.parametrize("K,D", [(512, 64), (544, 128), (1024, 256), (288, 512)])
def test_tc_operand_image_maps_are_injective(K, D):
    """Host-only: both layouts of the FP16 operand image (single CTA: 256-code blocks; CTA pair: half blocks, the lower /
    upper half of a chunk's codes for CTA 0 / 1) map every (code, column incl. the 16 fold columns) to its own 2-byte
    slot inside the image; in the pair layout the two halves of a chunk lie in different half blocks and a half block's
    codes are contiguous 16-byte rows per 8-column group (the K-major core-matrix rows the UMMA descriptor walks)."""
    from dvq import _cabi
    size = ctypes.c_longlong()
    for pair in (0, 1):
        seen = set()
        for k in range(K):
            for d in range(0, D + 16):
                off = _cabi.lib.dvq_debug_tc_image_offset(k, d, K, D, pair, ctypes.byref(size))
                assert 0 <= off and off + 2 <= size.value and off % 2 == 0, (pair, k, d, off, size.value)
                assert off not in seen, (pair, k, d)
                seen.add(off)
    off = lambda k, d, pair: _cabi.lib.dvq_debug_tc_image_offset(k, d, K, D, pair, None)
    # 8 consecutive columns of a code are one 16-byte row; the next code of the same half follows 16 bytes later
    assert [off(0, d, 1) - off(0, 0, 1) for d in range(8)] == [0, 2, 4, 6, 8, 10, 12, 14]
    assert off(1, 0, 1) - off(0, 0, 1) == 16 and off(1, 0, 0) - off(0, 0, 0) == 16
    # the last chunk's codes split in two halves (ragged K: half of what is left), each inside its own half block
    last = (K - 1) // 256 * 256
    nh = (min(256, K - last)) // 2
    ds = D if D <= 64 else (64 if D <= 256 else 32)
    half_bytes = (ds // 8 + 2) * 128 * 16
    assert off(last + nh, 0, 1) // half_bytes == off(last, 0, 1) // half_bytes + 1
    assert off(last + nh - 1, 0, 1) // half_bytes == off(last, 0, 1) // half_bytes
    assert _cabi.lib.dvq_debug_tc_image_offset(K, 0, K, D, 0, None) == -1


def test_module_caches_survive_deepcopy_and_see_reloaded_weights():
    """cached device-pointer structs made the modules unpicklable, and the fold / pack caches could go
    stale.  CPU-only checks of the host logic: deepcopy / pickle work, load_state_dict invalidates the caches."""
    import copy
    import pickle

    import torch
    import dvq
    from dvq.pixelcnn import GatedPixelCNN
    enc = dvq.PointNetEncoder(channel=4).eval()
    enc._folded = (torch.zeros(4), (0,))
    enc._folded_key = ("stale",)
    enc2 = copy.deepcopy(enc)
    assert enc2._folded is None and enc2._ws is None
    pickle.loads(pickle.dumps(enc))
    enc.load_state_dict(enc.state_dict())
    assert enc._folded is None                                # the post-hook dropped the cache
    enc._folded = (torch.zeros(4), (0,))
    enc.invalidate()
    assert enc._folded is None and enc.precision == "fp16_tc"
    vq = dvq.VectorQuantizer(16, 8, 0.25, 1.0)
    vq.process_group = object()
    vq2 = copy.deepcopy(vq)
    assert vq2.process_group is None and torch.equal(vq2.embedding.weight, vq.embedding.weight)
    pc = GatedPixelCNN(8, 8, 2, n_classes=4)
    packs, head = pc._pack()
    assert pc._packed is not None
    pc.load_state_dict(pc.state_dict())
    assert pc._packed is None
    copy.deepcopy(pc)

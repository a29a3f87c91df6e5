"""CPU oracle for the PointNet object encoder — TEST INFRASTRUCTURE ONLY.

numpy restatement of ``network/pointnet_encoder.py`` (``STN3d.forward`` :27-45,
``PointNetEncoder.forward`` :140-169, eval mode, ``global_feat=True``,
``feature_transform=False`` — the only configuration the reference constructs,
gen_net.py:16-17,31 / DVQVAE.py:18-19,33) and of ``utils.size_splits``
(utils/utils.py:163-181).  A checker, never a fallback: nothing under
``d-vqvae_b200/`` imports it.

Parity status: PINNED.  ``oracle/gen_golden.py`` loads the weights produced by
``make_state`` into the *real* reference module on CPU and stores its outputs
under ``tests/golden/pointnet_*.npz``; ``tests/test_oracle_golden.py`` checks
this restatement against them (FP32, tolerance 2e-5 of the feature scale: the
reference's conv/BN kernels and numpy's SGEMM sum in different orders).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
BN_EPS = 1e-5  # nn.BatchNorm1d default, pointnet_encoder.py:21-25,132-134


def _uniform(rs, shape, bound):
    return rs.uniform(-bound, bound, size=shape).astype(F32)


def _bn(rs, prefix, c, sd):
    sd[prefix + ".weight"] = rs.uniform(0.5, 1.5, size=c).astype(F32)
    sd[prefix + ".bias"] = (0.1 * rs.standard_normal(c)).astype(F32)
    sd[prefix + ".running_mean"] = (0.1 * rs.standard_normal(c)).astype(F32)
    sd[prefix + ".running_var"] = rs.uniform(0.5, 1.5, size=c).astype(F32)
    sd[prefix + ".num_batches_tracked"] = np.array(7, dtype=np.int64)


def make_state(seed: int, channel: int) -> dict:
    """Deterministic weights keyed exactly like the reference's ``state_dict()``
    (conv weights keep the trailing kernel dim of 1).  Non-trivial BN statistics
    so that the BN fold is actually exercised."""
    rs = np.random.RandomState(seed)
    sd = {}
    for pre in ("stn.", ""):
        for name, (co, ci) in (("conv1", (64, channel)), ("conv2", (128, 64)), ("conv3", (1024, 128))):
            b = 1.0 / np.sqrt(ci)
            sd[f"{pre}{name}.weight"] = _uniform(rs, (co, ci, 1), b)
            sd[f"{pre}{name}.bias"] = _uniform(rs, (co,), b)
        if pre:
            for name, (co, ci) in (("fc1", (512, 1024)), ("fc2", (256, 512)), ("fc3", (9, 256))):
                b = 1.0 / np.sqrt(ci)
                sd[f"stn.{name}.weight"] = _uniform(rs, (co, ci), b)
                sd[f"stn.{name}.bias"] = _uniform(rs, (co,), b)
        _bn(rs, pre + "bn1", 64, sd)
        _bn(rs, pre + "bn2", 128, sd)
        _bn(rs, pre + "bn3", 1024, sd)
        if pre:
            _bn(rs, "stn.bn4", 512, sd)
            _bn(rs, "stn.bn5", 256, sd)
    return sd


def make_cloud(seed: int, batch: int, channel: int, points: int) -> np.ndarray:
    """[B,C,P] fp32: xyz ~ 0.1*N(0,1); channel 3 (if present) is a per-cloud
    constant 'object scale' U(0.05,0.3) as in dataset_FHAB.py:50-54."""
    rs = np.random.RandomState(seed)
    x = (0.1 * rs.standard_normal((batch, channel, points))).astype(F32)
    if channel > 3:
        x[:, 3:, :] = rs.uniform(0.05, 0.3, size=(batch, channel - 3, 1)).astype(F32)
    return x


def _conv1x1(x, w, b):
    # nn.Conv1d(ci, co, 1) on [B,ci,P]: out[b,:,p] = W @ x[b,:,p] + bias
    return np.einsum("oc,bcp->bop", w[:, :, 0], x).astype(F32) + b[None, :, None]


def _bn_eval(x, sd, prefix):
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    inv = (F32(1.0) / np.sqrt(v + F32(BN_EPS))).astype(F32)
    shape = (1, -1, 1) if x.ndim == 3 else (1, -1)
    return ((x - m.reshape(shape)) * inv.reshape(shape) * g.reshape(shape) + b.reshape(shape)).astype(F32)


def _relu(x):
    return np.maximum(x, F32(0))


def stn3d_forward(x: np.ndarray, sd: dict, prefix: str = "stn.") -> np.ndarray:
    """pointnet_encoder.py:27-45."""
    h = _relu(_bn_eval(_conv1x1(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]), sd, prefix + "bn1"))
    h = _relu(_bn_eval(_conv1x1(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]), sd, prefix + "bn2"))
    h = _relu(_bn_eval(_conv1x1(h, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"]), sd, prefix + "bn3"))
    g = np.max(h, axis=2)                                                         # :32-33
    g = _relu(_bn_eval(g @ sd[prefix + "fc1.weight"].T + sd[prefix + "fc1.bias"], sd, prefix + "bn4"))
    g = _relu(_bn_eval(g @ sd[prefix + "fc2.weight"].T + sd[prefix + "fc2.bias"], sd, prefix + "bn5"))
    g = (g @ sd[prefix + "fc3.weight"].T + sd[prefix + "fc3.bias"]).astype(F32)
    iden = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], dtype=F32)[None, :]              # :39-43
    return (g + iden).astype(F32).reshape(-1, 3, 3)


def pointnet_forward(x: np.ndarray, sd: dict):
    """pointnet_encoder.py:140-166 — returns (feat [B,1024], trans [B,3,3], None)."""
    x = np.ascontiguousarray(x, dtype=F32)
    B, C, P = x.shape
    trans = stn3d_forward(x, sd)                                                  # :142
    xt = np.transpose(x, (0, 2, 1))                                               # :143
    xyz = np.matmul(xt[:, :, :3], trans).astype(F32)                              # :145-146 (size_splits + bmm)
    if C > 3:
        xt = np.concatenate([xyz, xt[:, :, 3:]], axis=2)                          # :148
    else:
        xt = xyz
    h = np.transpose(xt, (0, 2, 1))                                               # :149
    h = _relu(_bn_eval(_conv1x1(h, sd["conv1.weight"], sd["conv1.bias"]), sd, "bn1"))   # :150
    h = _relu(_bn_eval(_conv1x1(h, sd["conv2.weight"], sd["conv2.bias"]), sd, "bn2"))   # :161
    h = _bn_eval(_conv1x1(h, sd["conv3.weight"], sd["conv3.bias"]), sd, "bn3")          # :162 (no ReLU)
    feat = np.max(h, axis=2).astype(F32)                                          # :163-164
    return feat, trans, None

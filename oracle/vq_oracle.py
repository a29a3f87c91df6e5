"""CPU oracle for the VectorQuantizer hot path — TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the algorithm of the reference's
``network/vqvae/quantizer.py`` so the CUDA path can be checked on a machine
where ``/root/reference`` does not exist (the GPU box).  It is a checker, never
a fallback: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``d-vqvae_b200/`` imports it.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the *real* reference
(imported from /root/reference, module-global ``device`` re-pointed to CPU) on
seeded inputs and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those
fixtures (indices identical outside the documented FP32 near-tie band, z_q
bit-exact given the index, loss/perplexity within 1e-5 relative).

What can and cannot be bit-matched: the reference's
distance is ``d = fl(fl(zz + ee) - 2*dot)`` in FP32 with ``zz``, ``ee``, ``dot``
produced by torch reductions / SGEMM whose summation order is unspecified
(MKL on CPU, cuBLAS on GPU).  The *structure* and the first-minimum tie-break
are the contract; the low bits of ``d`` are not.  Hence the index-parity rule
``allowed_index_mismatch`` below (north_star: FP64 distance gap < 1e-6 rel).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
NEAR_TIE_REL = 1e-6  # north_star: "near-ties where the FP32 distance gap is below 1e-6 relative"


# --------------------------------------------------------------------------
# seeded synthetic inputs (numpy legacy RandomState: bit-stable across versions)
# --------------------------------------------------------------------------
def default_codebook(n_e: int, e_dim: int, seed: int) -> np.ndarray:
    """U(-1/n_e, 1/n_e) codebook — the reference's init distribution
    (quantizer.py:27), drawn from numpy so fixtures can be regenerated anywhere."""
    rs = np.random.RandomState(seed)
    return rs.uniform(-1.0 / n_e, 1.0 / n_e, size=(n_e, e_dim)).astype(F32)


def normal_latents(n: int, e_dim: int, seed: int) -> np.ndarray:
    rs = np.random.RandomState(seed)
    return rs.standard_normal(size=(n, e_dim)).astype(F32)


def variant_b(n: int, n_e: int, e_dim: int, seed: int):
    """Tie-free variant: E ~ N(0,1), z = E[randint] + 0.1*N(0,1)."""
    rs = np.random.RandomState(seed)
    E = rs.standard_normal(size=(n_e, e_dim)).astype(F32)
    pick = rs.randint(0, n_e, size=n)
    z = (E[pick] + F32(0.1) * rs.standard_normal(size=(n, e_dim)).astype(F32)).astype(F32)
    return z, E


# --------------------------------------------------------------------------
# the reference algorithm, FP32 structure
# --------------------------------------------------------------------------
def distances_f32(z: np.ndarray, E: np.ndarray) -> np.ndarray:
    """quantizer.py:36-38 / :46-48 —
    d = sum(z**2, dim=1, keepdim) + sum(E**2, dim=1) - 2 * z @ E.T, all FP32."""
    z = np.ascontiguousarray(z, dtype=F32)
    E = np.ascontiguousarray(E, dtype=F32)
    zz = np.sum(z * z, axis=1, keepdims=True, dtype=F32)          # [N,1]
    ee = np.sum(E * E, axis=1, dtype=F32)                          # [K]
    dot = z @ E.T                                                  # FP32 SGEMM
    return (zz + ee) - F32(2.0) * dot                              # fl(fl(zz+ee) - 2*dot)


def argmin_first(d: np.ndarray) -> np.ndarray:
    """quantizer.py:39 — torch.argmin(d, dim=1): lowest index among exact ties
    (verified).  np.argmin has the same first-occurrence rule."""
    return np.argmin(d, axis=1).astype(np.int64)


def one_hot(idx: np.ndarray, n_e: int) -> np.ndarray:
    """quantizer.py:40-42 — zeros(N, n_e).scatter_(1, idx, 1), FP32."""
    out = np.zeros((idx.shape[0], n_e), dtype=F32)
    out[np.arange(idx.shape[0]), idx.reshape(-1)] = 1.0
    return out


def zq_infer_from_idx(E: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """quantizer.py:43/53 — onehot @ E, which is bit-identical to the gather
    E[idx] (one non-zero term per output; SURVEY §0.5, verified)."""
    return np.ascontiguousarray(E[idx.reshape(-1)], dtype=F32)


def zq_train_from_idx(z: np.ndarray, E: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """quantizer.py:60 — z + (z_q - z).detach(): two FP32 roundings
    fl(z + fl(e - z)), no FMA contraction."""
    z2 = np.ascontiguousarray(z, dtype=F32).reshape(-1, E.shape[1])
    e = zq_infer_from_idx(E, idx)
    return (z2 + (e - z2).astype(F32)).astype(F32)


def loss_from_sse(sse: float, n_rows: int, e_dim: int, al: float, beta: float) -> np.float32:
    """quantizer.py:56-57 — al*mean((sg[z_q]-z)^2) + beta*mean((z_q-sg[z])^2).
    Forward value: both means are the same number m; torch evaluates
    fl32(al*m) + fl32(beta*m)."""
    m = F32(sse / (float(n_rows) * float(e_dim)))
    return F32(F32(F32(al) * m) + F32(F32(beta) * m))


def perplexity_from_hist(hist: np.ndarray, n_rows: int) -> np.float32:
    """quantizer.py:63-64 — e_mean = mean(onehot, 0); exp(-sum(e_mean*log(e_mean+1e-10)))."""
    p = hist.astype(np.float64) / float(n_rows)
    return F32(np.exp(-np.sum(p * np.log(p + 1e-10))))


def forward_infer(z: np.ndarray, E: np.ndarray):
    """quantizer.py:44-54 — returns (min_encoding_indices [N,1] int64, z_q like z)."""
    zf = np.ascontiguousarray(z, dtype=F32).reshape(-1, E.shape[1])
    idx = argmin_first(distances_f32(zf, E))
    return idx.reshape(-1, 1), zq_infer_from_idx(E, idx).reshape(z.shape)


def forward_train(z: np.ndarray, E: np.ndarray, al: float, beta: float, want_onehot: bool = True):
    """quantizer.py:30-43,56-67 — returns the reference's 5-tuple order
    (loss, z_q, perplexity, min_encodings, min_encoding_indices)."""
    zf = np.ascontiguousarray(z, dtype=F32).reshape(-1, E.shape[1])
    n, dim = zf.shape
    idx = argmin_first(distances_f32(zf, E))
    e = zq_infer_from_idx(E, idx)
    diff = (e - zf).astype(F32)
    sse = float(np.sum(diff.astype(np.float64) ** 2))
    loss = loss_from_sse(sse, n, dim, al, beta)
    zq = (zf + diff).astype(F32).reshape(z.shape)
    hist = np.bincount(idx, minlength=E.shape[0]).astype(np.int64)
    ppl = perplexity_from_hist(hist, n)
    enc = one_hot(idx, E.shape[0]) if want_onehot else None
    return loss, zq, ppl, enc, idx.reshape(-1, 1)


def get_emb(E: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """quantizer.py:68-75 — index -> embedding through a one-hot matmul; as
    written it only works for one index; the batched meaning is E[idx]."""
    return zq_infer_from_idx(E, np.asarray(idx).reshape(-1))


# --------------------------------------------------------------------------
# chunked form for sizes where N x K does not fit (per-row results
# are independent of row-chunking)
# --------------------------------------------------------------------------
def forward_stats_chunked(z: np.ndarray, E: np.ndarray, chunk: int = 65536):
    """idx [N], hist [K] int64, sse float64 — what the multi-GPU path all-reduces."""
    zf = np.ascontiguousarray(z, dtype=F32).reshape(-1, E.shape[1])
    n = zf.shape[0]
    idx = np.empty(n, dtype=np.int64)
    hist = np.zeros(E.shape[0], dtype=np.int64)
    sse = 0.0
    for s in range(0, n, chunk):
        zc = zf[s:s + chunk]
        ic = argmin_first(distances_f32(zc, E))
        idx[s:s + chunk] = ic
        hist += np.bincount(ic, minlength=E.shape[0])
        dc = (E[ic] - zc).astype(F32)
        sse += float(np.sum(dc.astype(np.float64) ** 2))
    return idx, hist, sse


# --------------------------------------------------------------------------
# FP64 truth + the index-parity rule
# --------------------------------------------------------------------------
def distances_f64_at(z: np.ndarray, E: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Exact-ish (FP64) squared distance of each row to the code idx[row]."""
    zf = np.asarray(z, dtype=np.float64).reshape(-1, E.shape[1])
    e = np.asarray(E, dtype=np.float64)[np.asarray(idx).reshape(-1)]
    return np.sum((zf - e) ** 2, axis=1)


def allowed_index_mismatch(z, E, idx_ours, idx_ref, rel: float = NEAR_TIE_REL):
    """SURVEY §8c rule.  Returns (n_mismatch, n_violations, worst_rel_excess).

    A mismatch at row n is allowed iff
        (d64[n, ours] - d64[n, ref]) / |d64[n, ref]| < rel
    i.e. our code is no farther than the reference's by more than the band
    (it may be *closer*: the reference's own FP32 rounding picks a non-optimal
    code on ~21/65536 rows of config 1)."""
    a = np.asarray(idx_ours).reshape(-1)
    b = np.asarray(idx_ref).reshape(-1)
    rows = np.nonzero(a != b)[0]
    if rows.size == 0:
        return 0, 0, 0.0
    zf = np.asarray(z).reshape(-1, E.shape[1])[rows]
    da = distances_f64_at(zf, E, a[rows])
    db = distances_f64_at(zf, E, b[rows])
    excess = (da - db) / np.maximum(np.abs(db), np.finfo(np.float64).tiny)
    return int(rows.size), int(np.sum(excess >= rel)), float(np.max(excess))


def truth_argmin_f64(z: np.ndarray, E: np.ndarray, chunk: int = 16384) -> np.ndarray:
    zf = np.asarray(z, dtype=np.float64).reshape(-1, E.shape[1])
    e = np.asarray(E, dtype=np.float64)
    ee = np.sum(e * e, axis=1)
    out = np.empty(zf.shape[0], dtype=np.int64)
    for s in range(0, zf.shape[0], chunk):
        zc = zf[s:s + chunk]
        d = np.sum(zc * zc, axis=1, keepdims=True) + ee - 2.0 * (zc @ e.T)
        out[s:s + chunk] = np.argmin(d, axis=1)
    return out

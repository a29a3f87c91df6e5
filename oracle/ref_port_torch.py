"""Torch-on-CPU port of the reference hot path — TEST/BENCH INFRASTRUCTURE ONLY.

The reference is a Python package that cannot travel to the GPU box
(/root/reference does not exist there), so the *CPU baseline* that
``bench.py`` reports (``cpu_baseline.kind == "port"``) and the
``--impl reference`` arm time this port: the same sequence of torch CPU ops the
reference issues (MKL SGEMM, ATen reductions, host one-hot, second SGEMM),
using every host thread torch will take.  ``oracle/gen_golden.py`` checks here,
where the real reference is importable, that this port returns bit-identical
tensors to ``network.vqvae.quantizer.VectorQuantizer.forward`` (same ops, same
library) — see tests/golden/port_vs_reference.json.

Not imported by anything under ``d-vqvae_b200/``.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def quantize_rows(z: torch.Tensor, codebook: torch.Tensor, al: float, beta: float, train: bool):
    """One call of the reference forward on a [N,D] CPU tensor
    (network/vqvae/quantizer.py:30-67), every intermediate materialised as the
    reference does: d [N,K], one-hot [N,K], one-hot @ E."""
    n_e, e_dim = codebook.shape
    flat = z.view(-1, e_dim)
    sq_z = torch.sum(flat ** 2, dim=1, keepdim=True)            # :36
    sq_e = torch.sum(codebook ** 2, dim=1)                      # :37
    d = sq_z + sq_e - 2 * torch.matmul(flat, codebook.t())      # :36-38
    nearest = torch.argmin(d, dim=1).unsqueeze(1)               # :39
    hot = torch.zeros(nearest.shape[0], n_e)                    # :40-41
    hot.scatter_(1, nearest, 1)                                 # :42
    z_q = torch.matmul(hot, codebook).view(z.shape)             # :43
    if not train:
        return nearest, z_q                                     # :54
    loss = al * torch.mean((z_q - z) ** 2) + beta * torch.mean((z_q - z) ** 2)   # :56-57
    z_q = z + (z_q - z)                                         # :60
    usage = torch.mean(hot, dim=0)                              # :63
    perplexity = torch.exp(-torch.sum(usage * torch.log(usage + 1e-10)))         # :64
    return loss, z_q, perplexity, hot, nearest                  # :67


@torch.no_grad()
def quantize_chunked_timed(z: torch.Tensor, codebook: torch.Tensor, al: float, beta: float, train: bool, chunk: int = 65536):
    """The timed form for bench.py: per 65 536-row chunk EXACTLY the reference forward (quantize_rows), and the
    cheapest possible combination of the scalar outputs — row-weighted mean of the chunk losses and an integer
    ``bincount`` of the chunk's indices for the perplexity.  No fp64 side computation is timed (round 1's
    ``quantize_chunked`` re-derived the histogram and the SSE in fp64, +36-44 % per chunk)."""
    n_e, e_dim = codebook.shape
    flat = z.view(-1, e_dim)
    n = flat.shape[0]
    idx = torch.empty(n, 1, dtype=torch.int64)
    z_q = torch.empty_like(flat)
    hist = torch.zeros(n_e, dtype=torch.int64)
    loss_rows = 0.0
    for s in range(0, n, chunk):
        zc = flat[s:s + chunk]
        if train:
            lc, qc, _, _, ic = quantize_rows(zc, codebook, al, beta, True)
            hist += torch.bincount(ic.view(-1), minlength=n_e)
            loss_rows += float(lc) * zc.shape[0]
        else:
            ic, qc = quantize_rows(zc, codebook, al, beta, False)
        idx[s:s + chunk] = ic
        z_q[s:s + chunk] = qc
    if not train:
        return idx, z_q.view(z.shape)
    p = hist.double() / n
    perplexity = torch.exp(-torch.sum(p * torch.log(p + 1e-10))).float()
    return torch.tensor(loss_rows / n, dtype=torch.float32), z_q.view(z.shape), perplexity, None, idx


@torch.no_grad()
def quantize_chunked(z: torch.Tensor, codebook: torch.Tensor, al: float, beta: float,
                     train: bool, chunk: int = 65536):
    """The reference cannot hold N x K at BASELINE config 2/4 sizes; per-row
    results are independent of row-chunking, so time it in
    65 536-row chunks and combine the scalar outputs from summed statistics."""
    n_e, e_dim = codebook.shape
    flat = z.view(-1, e_dim)
    n = flat.shape[0]
    idx = torch.empty(n, 1, dtype=torch.int64)
    z_q = torch.empty_like(flat)
    hist = torch.zeros(n_e, dtype=torch.float64)
    sse = 0.0
    for s in range(0, n, chunk):
        zc = flat[s:s + chunk]
        if train:
            _, qc, _, hot, ic = quantize_rows(zc, codebook, al, beta, True)
            hist += hot.sum(dim=0, dtype=torch.float64)
            sse += float(((codebook[ic.view(-1)] - zc).double() ** 2).sum())
        else:
            ic, qc = quantize_rows(zc, codebook, al, beta, False)
        idx[s:s + chunk] = ic
        z_q[s:s + chunk] = qc
    if not train:
        return idx, z_q.view(z.shape)
    m = torch.tensor(sse / (n * e_dim), dtype=torch.float32)
    loss = al * m + beta * m
    p = hist / n
    perplexity = torch.exp(-torch.sum(p * torch.log(p + 1e-10))).float()
    return loss, z_q.view(z.shape), perplexity, None, idx


def _bn_eval(x, w, b, mean, var, eps=1e-5):
    return torch.nn.functional.batch_norm(x, mean, var, w, b, False, 0.0, eps)


@torch.no_grad()
def pointnet_eval(x: torch.Tensor, sd: dict):
    """network/pointnet_encoder.py:27-45,140-166 with torch CPU ops (conv1d k=1,
    batch_norm eval, relu, max, linear, bmm) — the CPU baseline for the encoder."""
    F = torch.nn.functional

    def trunk(h, pre, last_relu):
        h = F.relu(_bn_eval(F.conv1d(h, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"]),
                            sd[pre + "bn1.weight"], sd[pre + "bn1.bias"], sd[pre + "bn1.running_mean"], sd[pre + "bn1.running_var"]))
        h = F.relu(_bn_eval(F.conv1d(h, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"]),
                            sd[pre + "bn2.weight"], sd[pre + "bn2.bias"], sd[pre + "bn2.running_mean"], sd[pre + "bn2.running_var"]))
        h = _bn_eval(F.conv1d(h, sd[pre + "conv3.weight"], sd[pre + "conv3.bias"]),
                     sd[pre + "bn3.weight"], sd[pre + "bn3.bias"], sd[pre + "bn3.running_mean"], sd[pre + "bn3.running_var"])
        if last_relu:
            h = F.relu(h)
        return torch.max(h, 2)[0]

    B, C, P = x.shape
    g = trunk(x, "stn.", True)
    g = F.relu(_bn_eval(F.linear(g, sd["stn.fc1.weight"], sd["stn.fc1.bias"]),
                        sd["stn.bn4.weight"], sd["stn.bn4.bias"], sd["stn.bn4.running_mean"], sd["stn.bn4.running_var"]))
    g = F.relu(_bn_eval(F.linear(g, sd["stn.fc2.weight"], sd["stn.fc2.bias"]),
                        sd["stn.bn5.weight"], sd["stn.bn5.bias"], sd["stn.bn5.running_mean"], sd["stn.bn5.running_var"]))
    g = F.linear(g, sd["stn.fc3.weight"], sd["stn.fc3.bias"])
    trans = (g + torch.eye(3, device=g.device).view(1, 9)).view(-1, 3, 3)
    xt = x.transpose(2, 1)
    xyz = torch.bmm(xt[:, :, :3], trans)
    xt = torch.cat([xyz, xt[:, :, 3:]], dim=2) if C > 3 else xyz
    feat = trunk(xt.transpose(2, 1).contiguous(), "", False)
    return feat, trans, None

"""Host-buffer (end-to-end) entry of the VQ path: ``dvq_vq_forward_host`` streams rows that live
in (pinned) HOST memory through the GPU in chunks — H2D copy, fused kernels and D2H copy of
consecutive chunks overlap on three streams inside the library — and returns host tensors.
This is the call ``bench.py`` times for its ``e2e`` number."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi


class HostQuantizer:
    def __init__(self, chunk_rows: int, n_e_max: int, e_dim_max: int, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.chunk_rows, self.n_e_max, self.e_dim_max = int(chunk_rows), int(n_e_max), int(e_dim_max)
        ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib.dvq_host_ctx_create(self.chunk_rows, self.n_e_max, self.e_dim_max, C.byref(ctx)),
                        "dvq_host_ctx_create")
        self._ctx = ctx

    def close(self):
        if self._ctx is not None and self._ctx.value:
            with torch.cuda.device(self.device):
                _cabi.lib.dvq_host_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, z: torch.Tensor, codebook: torch.Tensor, istrain: bool, al: float = 1.0, beta: float = 0.25,
                out_zq: torch.Tensor | None = None, out_idx: torch.Tensor | None = None, path: int = 0, copy_only: bool = False):
        """z [.., D] and codebook [K, D]: fp32 CPU tensors (pin them for full copy/compute overlap).
        Returns (loss, z_q, perplexity, idx[N,1]) if ``istrain`` else (idx[N,1], z_q) — host tensors."""
        if z.is_cuda or codebook.is_cuda:
            raise ValueError("HostQuantizer takes host tensors; use dvq.VectorQuantizer for device tensors")
        if z.dtype != torch.float32 or codebook.dtype != torch.float32 or not z.is_contiguous() or not codebook.is_contiguous():
            raise TypeError("z and codebook must be contiguous fp32")
        k, d = codebook.shape
        if z.numel() % d:
            raise RuntimeError("shape '[-1, %d]' is invalid for input of size %d" % (d, z.numel()))
        n = z.numel() // d
        zq = out_zq if out_zq is not None else torch.empty(z.shape, dtype=torch.float32, pin_memory=True)
        idx = out_idx if out_idx is not None else torch.empty((n, 1), dtype=torch.int64, pin_memory=True)
        loss, ppl = C.c_float(), C.c_float()
        flags = (_cabi.DVQ_TRAIN if istrain else 0) | path | (_cabi.DVQ_HOST_COPY_ONLY if copy_only else 0)   # copy_only: the pipeline's copies without kernels (bench ceiling; outputs undefined)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib.dvq_vq_forward_host(
                self._ctx, z.data_ptr(), codebook.data_ptr(), n, k, d, flags, float(al), float(beta),
                zq.data_ptr(), idx.data_ptr(), C.byref(loss), C.byref(ppl)), "dvq_vq_forward_host")
        if istrain:
            return loss.value, zq, ppl.value, idx
        return idx, zq

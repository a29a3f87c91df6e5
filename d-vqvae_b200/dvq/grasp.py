"""Batched grasp-generation graph around the two B200 hot-path modules — the inference path every
reference script exercises (``GenNet.gen``, network/gen_net.py:78-125), with the same sub-module names
(so a reference checkpoint's ``obj_encoder_*``, ``vqvae0..6``, ``decoder``, ``recon_encoder``,
``pos_decoder`` entries load unchanged) and *batched* semantics where the reference is B=1-only
(``idx6.repeat(1,3,3)`` / ``label = idx6[:,0,0]``, gen_net.py:88-89; ``get_emb``, quantizer.py:68-75).

Two things are pluggable because they are outside the hot path:
* ``prior(idx6 [B], batch) -> codes [B,6]`` int64 — the reference samples them with a GatedPixelCNN
  (gen_net.py:92-100); the default draws uniform codes (random-init benchmarks; the reference's own
  random-init PixelCNN emits out-of-range codes).  ``pixelcnn_prior(model)`` adapts a reference
  ``GatedPixelCNN`` instance.
* ``hand_layer(betas=[B,10], hand_pose=[B,45]) -> vertices [B,778,3]`` (or an object with ``.vertices``) — the reference
  uses the third-party MANO layer (gen_net.py:117-118): pass ``dvq.ManoLayer.from_pkl('models/mano/MANO_RIGHT.pkl')`` for the
  same arithmetic as one CUDA kernel; the default is a fixed linear stub (benchmarks without the MANO asset).
"""
from __future__ import annotations

import torch
import torch.nn as nn

import ctypes as C

from . import _cabi
from .pointnet import PointNetEncoder
from .vqvae import VQVAE


class Decoder(nn.Module):
    """MLP with the reference's parameter names ``MLP.L{i}.weight`` (network/DVQVAE.py:169-185)."""

    def __init__(self, layer_sizes, latent_size):
        super().__init__()
        self.MLP = nn.Sequential()
        widths = [latent_size] + list(layer_sizes)
        for i in range(len(layer_sizes)):
            self.MLP.add_module("L%d" % i, nn.Linear(widths[i], widths[i + 1]))
            if i + 1 < len(layer_sizes):
                self.MLP.add_module("A%d" % i, nn.ReLU())

    def forward(self, z):
        return self.MLP(z)


class LinearHandStub(nn.Module):
    """Deterministic stand-in for the MANO layer: template + fixed linear map of the 55 parameters."""

    def __init__(self, seed: int = 1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("template", 0.1 * torch.randn(778, 3, generator=g), persistent=False)
        self.register_buffer("basis", 0.01 * torch.randn(55, 778 * 3, generator=g), persistent=False)

    def forward(self, betas, hand_pose):
        return self.template + (torch.cat([betas, hand_pose], dim=1) @ self.basis).view(-1, 778, 3)


def uniform_prior(n_codes: int = 128, seed: int = 0):
    state = {}

    def prior(idx6, batch):
        gen = state.get(idx6.device)
        if gen is None:
            gen = state[idx6.device] = torch.Generator(device=idx6.device).manual_seed(seed)
        return torch.randint(0, n_codes, (batch, 6), device=idx6.device, generator=gen)
    return prior


def pixelcnn_prior(model, n_valid=None):
    """Adapter for a ``GatedPixelCNN`` (dvq.pixelcnn's row-cached sampler or the reference's): sample the 3x3
    grid conditioned on the object code and read the six hand-part positions (gen_net.py:92-100).  ``n_valid``
    (dvq sampler only) restricts the classes to the codebook size for random-init synthetic runs."""
    def prior(idx6, batch):
        grid = idx6.view(batch, 1, 1).repeat(1, 3, 3)
        kw = {} if n_valid is None else {"n_valid": n_valid}
        x = model.generate(grid, idx6.view(batch), shape=(3, 3), batch_size=batch, **kw)
        return torch.stack([x[:, 0, 1], x[:, 0, 2], x[:, 1, 1], x[:, 1, 2], x[:, 2, 1], x[:, 2, 2]], dim=1)
    return prior


class GraspGenerator(nn.Module):
    def __init__(self, prior=None, hand_layer=None):
        super().__init__()
        self.obj_encoder_type = PointNetEncoder(global_feat=True, feature_transform=False, channel=4)   # gen_net.py:16
        self.obj_encoder_pos = PointNetEncoder(global_feat=True, feature_transform=False, channel=4)    # :17
        for i in range(6):                                                                              # :20-25
            setattr(self, "vqvae%d" % i, VQVAE(128, 32, 2, 128, 256, 0.25, a=1))
        self.vqvae6 = VQVAE(128, 32, 2, 128, 1024, 2, a=0)                                              # :26
        self.decoder = Decoder([1024, 256, 55], 2560)                                                   # :27-28
        self.recon_encoder = PointNetEncoder(global_feat=True, feature_transform=False, channel=3)      # :31
        self.pos_decoder = Decoder([1024, 128, 6], 2048)                                                # :32-33
        self.prior = prior if prior is not None else uniform_prior(128)
        self.hand_layer = hand_layer if hand_layer is not None else LinearHandStub()
        self._graph = None

    @torch.no_grad()
    def gen_graphed(self, obj):
        """``gen`` replayed from a CUDA graph (captured on first use per input shape): the ~550 kernel launches of a
        batch — PointNets, VQ lookup, the PixelCNN sampler's GEMMs, softmax / multinomial, grouped gather, decoder MLPs —
        become one graph launch.  The prior must draw from torch's default CUDA generator (``pixelcnn_prior`` does);
        returns the graph's static output tensors (overwritten by the next call)."""
        key = (tuple(obj.shape), str(obj.device))
        if self._graph is None or self._graph[0] != key:
            static_in = obj.clone()
            side = torch.cuda.Stream(device=obj.device)
            side.wait_stream(torch.cuda.current_stream(obj.device))
            with torch.cuda.stream(side):
                for _ in range(2):                     # warm-up: workspaces, weight caches, kernel attributes
                    self.gen(static_in)
            torch.cuda.current_stream(obj.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.gen(static_in)
            self._graph = (key, graph, static_in, out)
        _, graph, static_in, out = self._graph
        static_in.copy_(obj)
        graph.replay()
        return out

    @torch.no_grad()
    def gen(self, obj):
        """obj [B,4,P] fp32 CUDA -> (recon [B,55], recon_pos [B,6]); stage tensors in ``self.last``."""
        B = obj.shape[0]
        feat_type, _, _ = self.obj_encoder_type(obj)                       # gen_net.py:81
        feat_pos, _, _ = self.obj_encoder_pos(obj)                         # :82
        idx6, obj_emb = self.vqvae6.inference(feat_type)                   # :83
        codes = self.prior(idx6.view(B), B)                                # :88-100
        # :101-106 + the torch.cat of :109 — six get_embbeding calls as ONE grouped gather that writes the decoder input
        z_out = torch.empty((B, 6 * 256 + 1024), dtype=torch.float32, device=obj.device)
        books = [getattr(self, "vqvae%d" % i).vector_quantization.embedding.weight.detach() for i in range(6)]
        ptrs = (C.c_void_p * 6)(*[w.data_ptr() for w in books])
        codes = codes.to(torch.int64).contiguous()
        oob = torch.zeros(1, dtype=torch.int32, device=obj.device)
        with torch.cuda.device(obj.device):
            _cabi.check(_cabi.lib.dvq_gather_multi(ptrs, 6, codes.data_ptr(), B, books[0].shape[0], 256, z_out.data_ptr(), z_out.shape[1],
                                                   oob.data_ptr(), torch.cuda.current_stream(obj.device).cuda_stream), "dvq_gather_multi")
        torch._assert_async(oob[0] == 0, "dvq.GraspGenerator: part-code index out of range")
        z_out[:, 6 * 256:] = feat_type
        recon = self.decoder(z_out).contiguous().view(B, 55)               # :109-113
        verts = self.hand_layer(betas=recon[:, :10], hand_pose=recon[:, 10:55])   # :117-118 (global_orient = transl = 0)
        verts = getattr(verts, "vertices", verts)                          # dvq.ManoLayer / mano / smplx return an output object
        hand_feat, _, _ = self.recon_encoder(verts.permute(0, 2, 1))       # :120
        recon_pos = self.pos_decoder(torch.cat([hand_feat, feat_pos], dim=1)).contiguous().view(B, 6)   # :121-123
        self.last = dict(feat_type=feat_type, feat_pos=feat_pos, idx6=idx6, obj_emb=obj_emb, codes=codes, verts=verts,
                         hand_feat=hand_feat)
        return recon, recon_pos

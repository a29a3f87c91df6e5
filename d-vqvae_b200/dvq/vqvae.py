"""``VQVAE`` wrapper — drop-in for the reference's ``network/VQVAE.py:11-53``: owns one
``VectorQuantizer`` as ``self.vector_quantization`` (so checkpoints keyed
``vqvaeN.vector_quantization.embedding.weight`` load unchanged, gen_diverse_grasp_obman.py:333-338)
and forwards to it.  The unused constructor arguments of the reference are kept."""
from __future__ import annotations

import torch.nn as nn

from .quantizer import VectorQuantizer


class VQVAE(nn.Module):
    def __init__(self, h_dim, res_h_dim, n_res_layers, n_embeddings, embedding_dim, beta, a=1,
                 save_img_embedding_map=False):
        super().__init__()
        self.vector_quantization = VectorQuantizer(n_embeddings, embedding_dim, beta, al=a)   # VQVAE.py:17-18
        self.img_to_embedding_map = {i: [] for i in range(n_embeddings)} if save_img_embedding_map else None

    def forward(self, inputs, verbose=False):
        """VQVAE.py:29-42 -> (embedding_loss, z_q, perplexity)."""
        embedding_loss, z_q, perplexity, _, _ = self.vector_quantization(inputs, True)
        if verbose:
            raise AssertionError("verbose=True is a debug trap in the reference (VQVAE.py:37-40)")
        return embedding_loss, z_q, perplexity

    def inference(self, inputs, verbose=False):
        """VQVAE.py:43-50 -> (min_encoding_indices [N,1] int64, z_q)."""
        return self.vector_quantization(inputs, False)

    def get_embbeding(self, index, dim):
        """VQVAE.py:51-53 (the reference's spelling is kept: gen_net.py:101-106 calls it)."""
        return self.vector_quantization.get_emb(index, dim)

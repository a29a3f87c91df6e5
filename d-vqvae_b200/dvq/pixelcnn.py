"""GatedPixelCNN prior of the grasp pipeline with an exact row-cached sampler.

Mirror of ``network/pixelcnn/models.py`` (``GatedMaskedConv2d`` :30-88, ``GatedPixelCNN`` :130-197):
same constructor, same sub-module names and parameter shapes (a reference ``state_dict`` loads
unchanged), same ``forward`` and the same sampling semantics in ``generate`` — including the
reference's quirks, which the sampler must reproduce to stay exact:

* the vertical stack of the mask-B layers sees the *current* row (kernel rows r-1, r), so the logits
  at (i, j) depend on the not-yet-sampled positions of row i, which hold index 0 at that time;
* ``probs = probs / probs.sum()`` normalises over the whole batch before ``multinomial``;
* the mask of the first layer is applied by zeroing the weights in place (``make_causal`` :61-63).

What the reference does per sampled position is a full forward over all 9 grid positions and 15
layers (:187-197).  Exactly the same logits need much less:

* the vertical stack at row r depends on rows <= r only, and rows < i are final when row i is being
  sampled -> their activations are cached for every layer; each step recomputes row i only (3 positions);
* the horizontal stack is needed at row i, columns <= j only (1-3 positions); the 1x1 output head at
  the single position (i, j).

(Row i-1 is refreshed once when row i starts: its last evaluation predates its final column.)
That is 33 + 18 + 9 position evaluations per 3x3 grid instead of 3 x 81: ~3.2x fewer FLOPs with
the same dependency structure (every GEMM is a dense [rows x taps*dim] x [taps*dim x 2dim]
product on the batch, which is what the tensor cores want at batch 4096).
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn as nn
import torch.nn.functional as F


class GatedMaskedConv2d(nn.Module):
    """models.py:30-88 (parameters only; the arithmetic lives in GatedPixelCNN)."""

    def __init__(self, mask_type, dim, kernel, residual=True, n_classes=128):
        super().__init__()
        assert kernel % 2 == 1, "Kernel size must be odd"
        self.mask_type = mask_type
        self.residual = residual
        self.kernel = kernel
        self.class_cond_embedding = nn.Embedding(n_classes, 2 * dim)
        self.vert_stack = nn.Conv2d(dim, dim * 2, (kernel // 2 + 1, kernel), 1, (kernel // 2, kernel // 2))
        self.vert_to_horiz = nn.Conv2d(2 * dim, 2 * dim, 1)
        self.horiz_stack = nn.Conv2d(dim, dim * 2, (1, kernel // 2 + 1), 1, (0, kernel // 2))
        self.horiz_resid = nn.Conv2d(dim, dim, 1)

    def make_causal(self):                                   # models.py:61-63
        self.vert_stack.weight.data[:, :, -1].zero_()
        self.horiz_stack.weight.data[:, :, :, -1].zero_()

    @staticmethod
    def gate(x):                                             # models.py:20-27
        a, b = x.chunk(2, dim=1)
        return torch.tanh(a) * torch.sigmoid(b)

    def forward(self, x_v, x_h, h):                          # models.py:65-88
        if self.mask_type == "A":
            self.make_causal()
        h = self.class_cond_embedding(h)
        h_vert = self.vert_stack(x_v)[:, :, :x_v.size(-1), :]
        out_v = self.gate(h_vert + h[:, :, None, None])
        h_horiz = self.horiz_stack(x_h)[:, :, :, :x_h.size(-2)]
        v2h = self.vert_to_horiz(h_vert)
        out = self.gate(v2h + h_horiz + h[:, :, None, None])
        out_h = self.horiz_resid(out) + x_h if self.residual else self.horiz_resid(out)
        return out_v, out_h


def _weights_init(m):                                        # models.py:9-16
    if m.__class__.__name__.find("Conv") != -1:
        try:
            nn.init.xavier_uniform_(m.weight.data)
            m.bias.data.fill_(0)
        except AttributeError:
            pass


def _invalidate_after_load(module, incompatible_keys):
    module.invalidate()


class GatedPixelCNN(nn.Module):
    def __init__(self, input_dim=256, dim=128, n_layers=15, n_classes=128):
        super().__init__()
        self.dim = dim
        self.input_dim = input_dim
        self.embedding = nn.Embedding(input_dim, dim)
        self.layers = nn.ModuleList()
        for i in range(n_layers):                            # models.py:141-149
            self.layers.append(GatedMaskedConv2d("A" if i == 0 else "B", dim, 5 if i == 0 else 3, i != 0, n_classes))
        self.output_conv = nn.Sequential(nn.Conv2d(dim, 2048, 1), nn.ReLU(True), nn.Conv2d(2048, input_dim, 1))
        self.apply(_weights_init)
        # the sampler's contractions: "fp32" / "tf32" = torch GEMMs (the reference's cuDNN convs are TF32 by default on a
        # GPU); "fp16_tc" = the repo's tcgen05 GEMM kernel with fused epilogues (pixelcnn_tc.py / csrc/pcnn_sm100.cu),
        # FP16 operands, FP32 accumulation (CUDA only, dim % 256 == 0, input_dim % 256 == 0)
        self.precision = "fp32"
        self._packed = None
        self._packed_tc = None
        self._tc_sampler = None
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    # ---- the reference forward, unchanged semantics (models.py:159-173) ---------------------------------
    def forward(self, x, label):
        shp = x.size() + (-1,)
        x = self.embedding(x.view(-1)).view(shp).permute(0, 3, 1, 2)
        x_v, x_h = x, x
        for layer in self.layers:
            x_v, x_h = layer(x_v, x_h, label)
        return self.output_conv(x_h)

    # ---- packed weights for the cached sampler ----------------------------------------------------------
    repack_every_call = False

    def invalidate(self):
        """Drop the packed weights.  The cache key is (data_ptr, _version) of every parameter: optimizer steps,
        ``copy_`` and ``.to()`` are seen; in-place edits through ``.data`` (the reference's ``weights_init`` /
        ``make_causal`` style) do not bump ``_version`` — call ``invalidate()`` after such an edit (a
        ``load_state_dict`` does it by itself), or set ``repack_every_call = True``."""
        self._packed = None
        self._packed_tc = None

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        self._packed_tc = None
        self._tc_sampler = None
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_packed"] = None
        d["_packed_tc"] = None
        d["_tc_sampler"] = None
        return d

    def _pack_tc(self):
        """FP16 weight images for the tcgen05 GEMM kernel; same cache key and invalidation as ``_pack``."""
        from . import pixelcnn_tc
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(next(self.parameters()).device),)
        if self._packed_tc is not None and self._packed_tc[0] == key and not self.repack_every_call:
            return self._packed_tc[1]
        self._packed_tc = (key, pixelcnn_tc.pack_weights(self))
        return self._packed_tc[1]

    def _pack(self):
        """[taps*dim, 2*dim] matrices, masks applied (as make_causal leaves the weights after the first
        reference forward).  Re-packed whenever a parameter version changes."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(next(self.parameters()).device),)
        if self._packed is not None and self._packed[0] == key and not self.repack_every_call:
            return self._packed[1]
        packs = []
        for layer in self.layers:
            k = layer.kernel
            wv = layer.vert_stack.weight.detach().clone()     # [2d, d, kh, kw]
            wh = layer.horiz_stack.weight.detach().clone()    # [2d, d, 1, kw2]
            if layer.mask_type == "A":
                wv[:, :, -1] = 0
                wh[:, :, :, -1] = 0
            kh = k // 2 + 1
            # one [d, taps * 2d] matrix per kernel row (column taps side by side): a grid row is multiplied once
            # and the per-tap results are shift-added, so no im2col copy is ever made
            wv_rows = [wv[:, :, a].permute(1, 2, 0).reshape(wv.shape[1], -1).contiguous() for a in range(kh)]   # [d, k*2d]
            packs.append(dict(
                k=k, residual=layer.residual, mask_a=layer.mask_type == "A",
                wv_rows=wv_rows, bv=layer.vert_stack.bias.detach(),
                wh=wh[:, :, 0].permute(1, 2, 0).reshape(wh.shape[1], -1).contiguous(),     # [d, kh*2d]
                bh=layer.horiz_stack.bias.detach(),
                wvh=layer.vert_to_horiz.weight.detach()[:, :, 0, 0].t().contiguous(), bvh=layer.vert_to_horiz.bias.detach(),
                wr=layer.horiz_resid.weight.detach()[:, :, 0, 0].t().contiguous(), br=layer.horiz_resid.bias.detach(),
                cond=layer.class_cond_embedding.weight.detach()))
        head = dict(w1=self.output_conv[0].weight.detach()[:, :, 0, 0].t().contiguous(), b1=self.output_conv[0].bias.detach(),
                    w2=self.output_conv[2].weight.detach()[:, :, 0, 0].t().contiguous(), b2=self.output_conv[2].bias.detach())
        self._packed = (key, (packs, head))
        return self._packed[1]

    def backend_name(self):
        """What executes the sampler's contractions (for benchmark records)."""
        return {"fp32": "FP32 cuBLAS GEMMs", "tf32": "TF32 cuBLAS GEMMs",
                "fp16_tc": "tcgen05 GEMM kernel with fused gate / residual epilogues (FP16 operands, FP32 accumulation)"}.get(self.precision, self.precision)

    def _matmul_ctx(self):
        if self.precision == "tf32" and torch.cuda.is_available():
            @contextlib.contextmanager
            def ctx():
                old = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = True
                try:
                    yield
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = old
            return ctx()
        return contextlib.nullcontext()

    @staticmethod
    def _gate(x, d):
        return torch.tanh(x[..., :d]) * torch.sigmoid(x[..., d:])

    @staticmethod
    def _shift_add(dst, y, taps, first_shift):
        """dst[:, c] += sum_b y[:, c + first_shift + b, b] over the columns that exist.  y: [B, Win, taps, C]."""
        wout, win = dst.shape[1], y.shape[1]
        for t in range(taps):
            sh = first_shift + t
            lo, hi = max(0, -sh), min(wout, win - sh)
            if hi > lo:
                dst[:, lo:hi] += y[:, lo + sh:hi + sh, t]

    def _above(self, packs, xv, r, cache):
        """Contribution of the finished rows (< r) to the vertical pre-activations of row r, per layer: constant
        while row r is being sampled."""
        above = []
        for l, pk in enumerate(packs):
            k, half = pk["k"], pk["k"] // 2
            src = xv[l]
            B, _, W, d = src.shape
            acc = pk["bv"].expand(B, W, -1).clone()
            for a in range(half):                            # kernel rows above the current one
                rr = r + a - half
                if rr >= 0:
                    y = (src[:, rr].reshape(B * W, d) @ pk["wv_rows"][a]).view(B, W, k, -1)
                    self._shift_add(acc, y, k, -half)
            above.append(acc)
        cache["above"] = above

    def _vert_row(self, packs, xv, x, label, r, cache):
        """Vertical stack of grid row r for every layer: cached part of the rows above + the current-row taps
        (mask-B layers).  Writes the gated activations into ``xv[l + 1][:, r]`` and returns the pre-activations
        (the horizontal stack's input)."""
        B, H, W = x.shape
        d = self.dim
        xv[0][:, r] = self.embedding(x[:, r, :])             # [B, W, d]  (x_v == x_h at the input, models.py:167)
        pre = []
        for l, pk in enumerate(packs):
            k, half = pk["k"], pk["k"] // 2
            h_vert = cache["above"][l]
            if not pk["mask_a"]:                             # the last kernel row (the current grid row) is masked in layer 0
                y = (xv[l][:, r].reshape(B * W, d) @ pk["wv_rows"][half]).view(B, W, k, -1)
                h_vert = h_vert.clone()
                self._shift_add(h_vert, y, k, -half)
            xv[l + 1][:, r] = self._gate(h_vert + pk["cond"][label][:, None, :], d)
            pre.append(h_vert)
        return pre

    @torch.no_grad()
    def step_logits(self, x, label, i, j, cache):
        """Logits of grid position (i, j) for the current index grid ``x`` [B, H, W] — equal to
        ``self.forward(x, label)[:, :, i, j]``.  ``cache`` (a dict owned by the caller) holds the
        vertical-stack activations of the finished rows; positions must be visited in raster order."""
        packs, head = self._pack()
        B, H, W = x.shape
        d = self.dim
        if "xv" not in cache:
            cache["xv"] = [x.new_zeros((B, H, W, d), dtype=self.embedding.weight.dtype) for _ in range(len(packs) + 1)]
        xv = cache["xv"]
        if j == 0:
            if i > 0:
                # row i-1 was last evaluated before its final column was sampled (the mask-B layers see the
                # whole current row): refresh it once with the finished indices, then it never changes again
                self._vert_row(packs, xv, x, label, i - 1, cache)
            self._above(packs, xv, i, cache)
        pre = self._vert_row(packs, xv, x, label, i, cache)
        xh = xv[0][:, i, : j + 1]                            # horizontal stack: row i, columns <= j
        n = j + 1
        for l, pk in enumerate(packs):
            half = pk["k"] // 2
            kh = half + 1
            cond = pk["cond"][label]                         # [B, 2d]
            taps = kh - 1 if pk["mask_a"] else kh            # layer 0: the current column is masked
            h_horiz = (pre[l][:, :n].reshape(B * n, -1) @ pk["wvh"]).view(B, n, -1) + (pk["bvh"] + pk["bh"])   # v2h + biases
            y = (xh.reshape(B * n, d) @ pk["wh"]).view(B, n, kh, -1)
            self._shift_add(h_horiz, y, taps, -half)         # taps at columns c-half..c
            out = self._gate(h_horiz + cond[:, None, :], d)
            res = (out.reshape(B * n, d) @ pk["wr"]).view(B, n, d) + pk["br"]
            xh = res + xh if pk["residual"] else res
        hid = torch.relu(xh[:, j] @ head["w1"] + head["b1"])
        return hid @ head["w2"] + head["b2"]                                      # [B, input_dim]

    @torch.no_grad()
    def generate(self, x_start, label, shape=(3, 3), batch_size=64, n_valid=None, forced=None, return_logits=False):
        """models.py:175-197 with the row-cached evaluation.  ``x_start`` is accepted and ignored exactly as
        in the reference.  ``n_valid``: keep only the first n classes (synthetic runs with random weights,
        where the codebooks have fewer rows than the PixelCNN has classes).  ``forced``: [B, H, W] indices to
        write instead of sampling (parity tests)."""
        param = next(self.parameters())
        x = torch.zeros((batch_size, *shape), dtype=torch.int64, device=param.device)
        cache, all_logits = {}, []
        tc = None
        if self.precision == "fp16_tc":
            from . import pixelcnn_tc
            tc = self._tc_sampler
            if tc is None or tc.B != batch_size or (tc.H, tc.W) != tuple(shape) or tc.dev != param.device:
                tc = self._tc_sampler = pixelcnn_tc.TcSampler(self, batch_size, tuple(shape))
            tc.begin(label)
        elif self.precision not in ("fp32", "tf32"):
            raise ValueError("precision must be 'fp32', 'tf32' or 'fp16_tc'")
        with self._matmul_ctx():
            for i in range(shape[0]):
                for j in range(shape[1]):
                    logits = tc.step_logits(x, i, j) if tc is not None else self.step_logits(x, label, i, j, cache)
                    if return_logits:
                        all_logits.append(logits.clone() if tc is not None else logits)
                    if n_valid is not None:
                        logits = logits.clone()
                        logits[:, n_valid:] = float("-inf")
                    probs = F.softmax(logits, -1)
                    probs = probs / probs.sum()                                   # models.py:194 (batch-wide, kept)
                    if forced is not None:
                        x[:, i, j] = forced[:, i, j]
                    else:
                        x[:, i, j] = probs.multinomial(1).squeeze(-1)
        return (x, all_logits) if return_logits else x

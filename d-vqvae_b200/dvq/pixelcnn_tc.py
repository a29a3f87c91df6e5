"""Kernel-backed evaluation of the GatedPixelCNN row-cached sampler (``GatedPixelCNN.precision = "fp16_tc"``).

Same algorithm as ``GatedPixelCNN.step_logits`` (row-cached, exact dependency structure of
network/pixelcnn/models.py:65-88,176-197), but every contraction is one launch of the tcgen05 GEMM with a fused
epilogue in ``csrc/pcnn_sm100.cu`` (``dvq_pcnn_gemm``): FP16 operands, FP32 accumulation, activations kept in HBM
as FP16 operand images that the next launch consumes directly.  Per sampled position (i, j), per layer:

* vertical stack of grid row i (all W columns): ONE launch — a K-segment per kernel tap (rows above + the current
  row, column offsets as whole-tile shifts), epilogue adds bias + class-conditional row, writes the gated activation
  (input of the next layer's vertical stack) and the raw pre-activation (input of the vertical-to-horizontal conv);
* horizontal stack, columns <= j: ONE launch for ``gate(v2h(pre) + horiz taps + bias + cond)`` and ONE for the 1x1
  residual convolution (+ FP32 residual stream);
* the 1x1 output head at column j: two launches (ReLU hidden layer, logits).

Host work per launch is a ctypes call with a descriptor struct; softmax / multinomial stay in torch (9 tiny ops per
sample grid).  No CPU or library-GEMM fallback: ``precision = "fp16_tc"`` requires CUDA tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi

PC_GATE, PC_RES, PC_RELU, PC_LOGITS = 0, 1, 2, 3


class _Seg(C.Structure):
    _fields_ = [("a_img", C.c_void_p), ("w_img", C.c_void_p), ("a_kd", C.c_int), ("ks", C.c_int),
                ("tile_shift", C.c_int), ("col_shift", C.c_int)]


class _Gemm(C.Structure):
    _fields_ = [("seg", _Seg * 12), ("nseg", C.c_int), ("m_tiles", C.c_int), ("n_tiles", C.c_int), ("tiles_per_col", C.c_int),
                ("ncols_src", C.c_int), ("mode", C.c_int), ("bias", C.c_void_p), ("cond_img", C.c_void_p), ("out_img", C.c_void_p),
                ("pre_img", C.c_void_p), ("res_img", C.c_void_p), ("logits", C.c_void_p), ("out_kd", C.c_int), ("res_in", C.c_int),
                ("d_gate", C.c_int), ("err", C.c_void_p)]


def _w_image(wm: torch.Tensor, perm=None) -> torch.Tensor:
    """[K, N] fp32 matrix -> FP16 weight image [N/256][K/8][256][8] (columns optionally permuted first)."""
    if perm is not None:
        wm = wm[:, perm]
    k, n = wm.shape
    assert k % 8 == 0 and n % 256 == 0, (k, n)
    return wm.reshape(k // 8, 8, n // 256, 256).permute(2, 0, 3, 1).contiguous().to(torch.float16)


def gate_perm(d: int, device) -> torch.Tensor:
    """Column order of a gated GEMM: output tile t holds the 'a' columns 128 t .. 128 t + 127 and the matching 'b' columns."""
    t = torch.arange(d // 128, device=device).view(-1, 1) * 128
    a = t + torch.arange(128, device=device).view(1, -1)
    return torch.cat([a, a + d], dim=1).reshape(-1)


def pack_weights(model):
    """Weight images / bias vectors of every contraction of the sampler (masks applied as make_causal leaves them)."""
    d = model.dim
    dev = model.embedding.weight.device
    if d % 256 or model.input_dim % 256:
        raise ValueError("precision='fp16_tc' needs dim %% 256 == 0 and input_dim %% 256 == 0 (got %d, %d)" % (d, model.input_dim))
    perm = gate_perm(d, dev)
    layers = []
    for layer in model.layers:
        k = layer.kernel
        half, kh = k // 2, k // 2 + 1
        wv = layer.vert_stack.weight.detach().float()      # [2d, d, kh, k]
        wh = layer.horiz_stack.weight.detach().float()     # [2d, d, 1, kh]
        mask_a = layer.mask_type == "A"
        vert = {}
        for a in range(kh):
            if mask_a and a == kh - 1:
                continue                                   # the current grid row is masked in layer 0 (models.py:61-62)
            for t in range(k):
                vert[(a, t)] = _w_image(wv[:, :, a, t].t(), perm)
        horiz = {}
        for t in range(kh):
            if mask_a and t == kh - 1:
                continue                                   # the current column is masked in layer 0 (:63)
            horiz[t] = _w_image(wh[:, :, 0, t].t(), perm)
        layers.append(dict(
            k=k, half=half, kh=kh, residual=bool(layer.residual), vert=vert, horiz=horiz,
            bv=layer.vert_stack.bias.detach().float()[perm].contiguous(),
            wvh=_w_image(layer.vert_to_horiz.weight.detach().float()[:, :, 0, 0].t(), perm),
            bh=(layer.vert_to_horiz.bias.detach().float() + layer.horiz_stack.bias.detach().float())[perm].contiguous(),
            wr=_w_image(layer.horiz_resid.weight.detach().float()[:, :, 0, 0].t()),
            br=layer.horiz_resid.bias.detach().float().contiguous(),
            cond=layer.class_cond_embedding.weight.detach().float().contiguous()))
    head = dict(w1=_w_image(model.output_conv[0].weight.detach().float()[:, :, 0, 0].t()), b1=model.output_conv[0].bias.detach().float().contiguous(),
                w2=_w_image(model.output_conv[2].weight.detach().float()[:, :, 0, 0].t()), b2=model.output_conv[2].bias.detach().float().contiguous())
    zero_w = torch.zeros((2 * d // 256) * 8 * 256 * 8, dtype=torch.float16, device=dev)   # K = 64 of zeros, any gated N
    return dict(layers=layers, head=head, zero_w=zero_w, emb=model.embedding.weight.detach().float().contiguous())


class TcSampler:
    """Buffers + launch sequence for one (batch size, grid shape)."""

    def __init__(self, model, batch: int, shape):
        self.m = model
        self.B, (self.H, self.W) = int(batch), shape
        self.Bp = (self.B + 127) // 128 * 128
        self.T = self.Bp // 128
        d, L = model.dim, len(model.layers)
        dev = model.embedding.weight.device
        if dev.type != "cuda":
            raise ValueError("precision='fp16_tc' has no CPU path: move the model to a CUDA device")
        self.dev, self.d, self.L = dev, d, L
        self.packs = None
        rows = self.W * self.Bp
        h16 = lambda n: torch.empty(n, dtype=torch.float16, device=dev)   # every image is written in full before it is read
        self.xv = [[h16(rows * d) for _ in range(self.H)] for _ in range(L + 1)]
        self.pre = [h16(rows * 2 * d) for _ in range(L)]
        self.xh16 = h16(rows * d)
        self.out16 = h16(rows * d)
        self.xh32 = torch.empty(rows * d, dtype=torch.float32, device=dev)
        self.hid16 = h16(self.Bp * 2048)
        self.logits = torch.empty((self.Bp, model.input_dim), dtype=torch.float32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cond = [h16(self.Bp * 2 * d) for _ in range(L)]
        self.flops = 0          # 2 * M * K * N issued to the tensor cores since begin() (padded rows included)

    def begin(self, label: torch.Tensor):
        """Start a new sample grid: current weight images, class-conditional rows of this batch's labels."""
        self.flops = 0
        self.packs = self.m._pack_tc()
        label = label.to(self.dev).to(torch.int64).contiguous()
        with torch.cuda.device(self.dev):
            for l in range(self.L):
                tab = self.packs["layers"][l]["cond"]
                _cabi.check(_cabi.lib.dvq_pcnn_rows_to_image(label.data_ptr(), self.B, self.Bp, tab.data_ptr(), tab.shape[0], 2 * self.d,
                                                             self.cond[l].data_ptr(), self._stream()), "dvq_pcnn_rows_to_image")

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _gemm(self, segs, m_tiles, n_tiles, ncols_src, mode, bias, cond=None, out=None, pre=None, res=None, res_in=0, logits=None, out_kd=0):
        g = _Gemm()
        assert 0 < len(segs) <= 12
        for i, (a_ptr, w, a_kd, ks, col_shift) in enumerate(segs):
            g.seg[i].a_img, g.seg[i].w_img = a_ptr, w.data_ptr()
            g.seg[i].a_kd, g.seg[i].ks, g.seg[i].tile_shift, g.seg[i].col_shift = a_kd, ks, col_shift * self.T, col_shift
        g.nseg, g.m_tiles, g.n_tiles, g.tiles_per_col, g.ncols_src, g.mode = len(segs), m_tiles, n_tiles, self.T, ncols_src, mode
        for c in range(m_tiles // self.T):       # accounting: the segments the kernel runs for the tiles of grid column c
            k_valid = sum(ks for (_, _, _, ks, sh) in segs if 0 <= c + sh < ncols_src)
            self.flops += 2 * (self.T * 128) * k_valid * (n_tiles * 256)
        g.bias = bias.data_ptr()
        g.cond_img = cond.data_ptr() if cond is not None else None
        g.out_img = out.data_ptr() if out is not None else None
        g.pre_img = pre.data_ptr() if pre is not None else None
        g.res_img = res.data_ptr() if res is not None else None
        g.logits = logits.data_ptr() if logits is not None else None
        g.out_kd, g.res_in, g.d_gate, g.err = out_kd, res_in, self.d, self.err.data_ptr()
        _cabi.check(_cabi.lib.dvq_pcnn_gemm(C.byref(g), self._stream()), "dvq_pcnn_gemm")

    def _embed_row(self, x, r, img16, img32=None):
        row = x[:, r, :].contiguous()
        emb = self.packs["emb"]
        _cabi.check(_cabi.lib.dvq_pcnn_embed(row.data_ptr(), self.W, self.W, self.B, self.Bp, emb.data_ptr(), emb.shape[0], self.d,
                                             img16.data_ptr(), img32.data_ptr() if img32 is not None else None, self._stream()), "dvq_pcnn_embed")

    def _vertical_row(self, r):
        """Vertical stack of grid row r for every layer (inputs xv[l][rows <= r], outputs xv[l + 1][r] and pre[l])."""
        d = self.d
        for l, pk in enumerate(self.packs["layers"]):
            segs = []
            for (a, t), w in pk["vert"].items():
                rr = r + a - pk["half"]
                if rr < 0:
                    continue
                segs.append((self.xv[l][rr].data_ptr(), w, d, d, t - pk["half"]))
            if not segs:       # layer 0 on the first grid row: nothing above, current row masked -> gate(bias + cond)
                segs.append((self.xv[l][r].data_ptr(), self.packs["zero_w"], d, 64, 0))
            self._gemm(segs, self.W * self.T, d // 128, self.W, PC_GATE, pk["bv"], cond=self.cond[l], out=self.xv[l + 1][r], pre=self.pre[l], out_kd=d)

    def step_logits(self, x, i, j):
        """Logits [B, input_dim] of grid position (i, j) for the current index grid x [B, H, W] (raster order)."""
        d, T = self.d, self.T
        with torch.cuda.device(self.dev):
            if j == 0 and i > 0:
                # row i-1 was last evaluated before its final column was sampled: refresh it once, then it never changes
                self._embed_row(x, i - 1, self.xv[0][i - 1])
                self._vertical_row(i - 1)
            self._embed_row(x, i, self.xv[0][i])
            self._embed_row(x, i, self.xh16, self.xh32)
            self._vertical_row(i)
            n = j + 1
            for l, pk in enumerate(self.packs["layers"]):
                segs = [(self.pre[l].data_ptr(), pk["wvh"], 2 * d, 2 * d, 0)]
                for t, w in pk["horiz"].items():
                    segs.append((self.xh16.data_ptr(), w, d, d, t - pk["half"]))
                self._gemm(segs, n * T, d // 128, n, PC_GATE, pk["bh"], cond=self.cond[l], out=self.out16, out_kd=d)
                self._gemm([(self.out16.data_ptr(), pk["wr"], d, d, 0)], n * T, d // 256, n, PC_RES, pk["br"], out=self.xh16, res=self.xh32,
                           res_in=1 if pk["residual"] else 0, out_kd=d)
            hd = self.packs["head"]
            col_ptr = self.xh16.data_ptr() + j * T * d * 256                     # the tiles of grid column j
            self._gemm([(col_ptr, hd["w1"], d, d, 0)], T, 2048 // 256, 1, PC_RELU, hd["b1"], out=self.hid16, out_kd=2048)
            self._gemm([(self.hid16.data_ptr(), hd["w2"], 2048, 2048, 0)], T, self.m.input_dim // 256, 1, PC_LOGITS, hd["b2"], logits=self.logits)
        return self.logits[:self.B]

    def check(self):
        """Synchronous read of the pipeline error word (tests / debugging)."""
        e = int(self.err.item())
        if e:
            raise RuntimeError("pcnn_gemm_kernel pipeline time-out (code %d)" % e)

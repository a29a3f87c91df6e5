"""``PointNetEncoder`` / ``STN3d`` — drop-in for ``network/pointnet_encoder.py:10-45,125-169``.

Same constructor ``(global_feat=True, feature_transform=False, channel=3)``, same sub-module
and parameter names (``stn.conv1..3``, ``stn.fc1..3``, ``stn.bn1..5``, ``conv1..3``, ``bn1..3``
incl. ``running_mean/var/num_batches_tracked``) so reference checkpoints load with
``load_state_dict`` unchanged; ``forward(x [B,C,P]) -> (feat [B,1024], trans [B,3,3], None)``.

The forward is ONE call into the C ABI (``dvq_pointnet_forward``): eval-mode BatchNorm is folded
into the preceding conv / linear here on the host side (cached until a parameter or running
statistic changes), the 3x3 input transform is fused into the point load and no activation
ever reaches HBM.  Only what the reference constructs is implemented: ``global_feat=True``,
``feature_transform=False`` (gen_net.py:16-17,31; DVQVAE.py:18-20,33), eval mode.  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _cabi


class STN3d(nn.Module):
    """Parameter container mirroring pointnet_encoder.py:11-25 (its forward runs inside the fused kernel)."""

    def __init__(self, channel):
        super().__init__()
        self.conv1 = nn.Conv1d(channel, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, 9)
        self.relu = nn.ReLU()
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.bn4 = nn.BatchNorm1d(512)
        self.bn5 = nn.BatchNorm1d(256)

    def forward(self, x):
        raise RuntimeError("dvq.STN3d is evaluated inside dvq.PointNetEncoder's fused kernel; call the encoder")


def _fold(lin_w, lin_b, bn):
    """Fold eval-mode BatchNorm into the preceding 1x1 conv / linear (fp64 on device, cast to fp32)."""
    w = lin_w.detach().double().reshape(lin_w.shape[0], -1)
    b = lin_b.detach().double()
    if bn is None:
        return w.float().contiguous(), b.float().contiguous()
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    return (w * s[:, None]).float().contiguous(), ((b - bn.running_mean.detach().double()) * s + bn.bias.detach().double()).float().contiguous()


def _invalidate_after_load(module, incompatible_keys):
    module.invalidate()


class PointNetEncoder(nn.Module):
    def __init__(self, global_feat=True, feature_transform=False, channel=3):
        super().__init__()
        if not global_feat or feature_transform:
            raise NotImplementedError(
                "dvq.PointNetEncoder implements the configuration the reference constructs "
                "(global_feat=True, feature_transform=False: gen_net.py:16-17,31)")
        if channel not in (3, 4):
            raise NotImplementedError("channel must be 3 (hand vertices) or 4 (xyz + object scale)")
        self.stn = STN3d(channel)
        self.conv1 = nn.Conv1d(channel, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.global_feat = global_feat
        self.feature_transform = feature_transform
        self.channel = channel
        # "fp16_tc" (default): layers 2-3 on tcgen05 tensor cores, FP16 operands / FP32 accumulation — the same 10-bit
        # operand mantissa as the TF32 cuDNN convolutions the reference runs on a GPU (torch.backends.cudnn.allow_tf32
        # is True by default), parity 1e-3 of the feature scale; "fp32": all-FP32 CUDA-core kernel (parity 5e-5 vs the
        # CPU reference, 14x slower)
        self.precision = "fp16_tc"
        self._folded = None
        self._folded_key = None
        self._ws = None
        # a checkpoint load replaces parameter contents without necessarily bumping every version counter
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    # -------------------------------------------------------------- BN fold cache
    def invalidate(self):
        """Drop the cached BN-folded weights.  The cache is keyed on (data_ptr, _version, device) of every
        parameter and buffer, which catches optimizer steps, ``copy_``, ``load_state_dict`` and ``.to()``, but NOT
        in-place edits through ``.data`` (``w.data.mul_()``, ``running_mean.data.fill_()`` do not bump ``_version``):
        after such an edit call ``invalidate()``, or set ``refold_every_call = True``."""
        self._folded = None
        self._folded_key = None

    refold_every_call = False
    _warned_detached = False

    def _apply(self, fn, *args, **kwargs):     # .to() / .cuda() / .float(): new storage
        self.invalidate()
        self._ws = None
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):                    # deepcopy / pickle / torch.save(module): caches hold raw device pointers
        d = self.__dict__.copy()
        d["_folded"] = None
        d["_folded_key"] = None
        d["_ws"] = None
        return d

    def _fold_key(self):
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in list(self.parameters()) + list(self.buffers()))

    def folded_weights(self):
        key = self._fold_key()
        if self._folded is not None and key == self._folded_key and not self.refold_every_call:
            blob, offs = self._folded
            return blob, self._weights_struct(blob, offs)
        s = self.stn
        parts = [
            _fold(s.conv1.weight, s.conv1.bias, s.bn1), _fold(s.conv2.weight, s.conv2.bias, s.bn2),
            _fold(s.conv3.weight, s.conv3.bias, s.bn3), _fold(s.fc1.weight, s.fc1.bias, s.bn4),
            _fold(s.fc2.weight, s.fc2.bias, s.bn5), _fold(s.fc3.weight, s.fc3.bias, None),
            _fold(self.conv1.weight, self.conv1.bias, self.bn1), _fold(self.conv2.weight, self.conv2.bias, self.bn2),
            _fold(self.conv3.weight, self.conv3.bias, self.bn3),
        ]
        flat = [t for pair in parts for t in pair]
        # one allocation, every tensor 16-byte aligned (128-bit weight loads)
        offs, total = [], 0
        for t in flat:
            offs.append(total)
            total += (t.numel() + 3) // 4 * 4
        blob = torch.zeros(total, dtype=torch.float32, device=flat[0].device)
        for t, o in zip(flat, offs):
            blob[o:o + t.numel()] = t.reshape(-1)
        self._folded = (blob, tuple(offs))     # (only tensors and ints are cached: the module stays picklable)
        self._folded_key = key
        return blob, self._weights_struct(blob, offs)

    @staticmethod
    def _weights_struct(blob, offs):
        st = _cabi.PointNetWeights()
        for (name, _), o in zip(_cabi.PointNetWeights._fields_, offs):
            setattr(st, name, blob.data_ptr() + 4 * o)
        return st

    def forward(self, x):
        if self.training:
            raise RuntimeError("dvq.PointNetEncoder fuses eval-mode BatchNorm; call .eval() first "
                               "(every reference script does: gen_diverse_grasp_obman.py:347)")
        if not isinstance(x, torch.Tensor) or x.dtype != torch.float32:
            raise TypeError("x must be an fp32 tensor [B, C, P]")
        if x.dim() != 3 or x.shape[1] != self.channel:
            raise ValueError("x must be [B, %d, P]; got %s" % (self.channel, tuple(x.shape)))
        if not x.is_cuda or not self.conv1.weight.is_cuda:
            raise ValueError("dvq.PointNetEncoder has no CPU path: x and the weights must be CUDA tensors")
        if torch.is_grad_enabled():
            # the fused kernel has no backward.  An input that asks for gradients is an error; parameters that merely
            # still have requires_grad=True (the reference scripts call GenNet.gen without torch.no_grad():
            # gen_diverse_grasp_obman.py:236) get detached outputs and a one-time warning
            if x.requires_grad:
                raise RuntimeError("dvq.PointNetEncoder is inference-only (no autograd through the fused kernel): "
                                   "detach x or call the encoder under torch.no_grad()")
            if not PointNetEncoder._warned_detached and any(p.requires_grad for p in self.parameters()):
                PointNetEncoder._warned_detached = True
                import warnings
                warnings.warn("dvq.PointNetEncoder returns detached features: its parameters will receive no gradient "
                              "(inference-only fused kernel); use the reference encoder for training / fine-tuning")
        x = x.detach()
        if not x.is_contiguous():
            x = x.contiguous()   # e.g. the permute at gen_net.py:120
        B, Cc, P = x.shape
        blob, st = self.folded_weights()
        feat = torch.empty((B, 1024), dtype=torch.float32, device=x.device)
        trans = torch.empty((B, 3, 3), dtype=torch.float32, device=x.device)
        if self.precision not in ("fp32", "fp16_tc"):
            raise ValueError("precision must be 'fp32' or 'fp16_tc'")
        flags = _cabi.DVQ_PN_FP16_TC if self.precision == "fp16_tc" else 0
        need = _cabi.pointnet_workspace_bytes(B, Cc, P, flags)
        if self._ws is None or self._ws.device != x.device or self._ws.numel() < need:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib.dvq_pointnet_forward_ex(
                x.data_ptr(), C.addressof(st), B, Cc, P, flags, feat.data_ptr(), trans.data_ptr(),
                self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream(x.device).cuda_stream),
                "dvq_pointnet_forward_ex")
        return feat, trans, None

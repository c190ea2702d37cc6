"""Inference-only evaluation of a FoldingNet decoder for the latent-grid scans.

The decoder is not a kernel target of this project (its convolutions run on
cuDNN), but as the reference writes it - ``Conv1d -> BatchNorm1d -> ReLU`` on NCHW tensors
(src/molearn/models/foldingnet.py:119-153) - two thirds of a decode batch on a B200 are memory-bound glue
around the convolutions (torch profiler, batch 256, MurD: BatchNorm 19 %, bias add 20 %, ReLU 8 %,
NCHW<->NHWC conversions inside cuDNN 19 %; the convolutions themselves 31 %).  ``FusedDecoder`` evaluates
the SAME network (any module with the reference's attribute layout: ``layer1..4`` -> ``fold1/fold2`` ->
``layers``; weights are read, never modified) with

  * eval-mode BatchNorm folded into the preceding convolution's weights and bias,
  * bias + ReLU fused into the convolution (``torch.cudnn_convolution_relu``),
  * channels-last activations end to end (no layout conversions), channel counts padded to multiples of 8
    with zero weights (exact),

and returns the ``[B, M, 3]`` point cloud directly (the reference returns a permuted view of ``[B, 3, M]``;
the energy / RMSD kernels take any strides).  The arithmetic is the reference's up to the rounding of the
folded weights (1e-6 relative) and cuDNN's TF32 convolution math, which the reference's own decode uses too.
It is a snapshot: rebuild it after the network's weights change.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def _round_up(n, k=8):
    return (n + k - 1) // k * k


class _Fold:
    """One FoldingLayer: list of (weight [O_pad, I_pad, 1, 3] channels-last, bias [O_pad], relu, O)."""

    def __init__(self, folding_layer, device):
        mods = list(folding_layer.layers)
        self.convs = []
        i = 0
        in_pad = None
        while i < len(mods):
            conv = mods[i]
            if not isinstance(conv, nn.Conv1d) or conv.kernel_size != (3,) or conv.padding != (1,) or conv.stride != (1,):
                raise TypeError(f"unsupported layer in a folding stack: {conv}")
            w = conv.weight.detach().to(device=device, dtype=torch.float32)
            b = (conv.bias.detach().to(device=device, dtype=torch.float32) if conv.bias is not None
                 else torch.zeros(w.shape[0], device=device))
            i += 1
            relu = False
            if i < len(mods) and isinstance(mods[i], nn.BatchNorm1d):
                bn = mods[i]
                scale = (bn.weight.detach().to(device).float() if bn.affine else torch.ones_like(b)) / torch.sqrt(
                    bn.running_var.detach().to(device).float() + bn.eps)
                shift = bn.bias.detach().to(device).float() if bn.affine else torch.zeros_like(b)
                w = w * scale[:, None, None]
                b = (b - bn.running_mean.detach().to(device).float()) * scale + shift
                i += 1
            if i < len(mods) and isinstance(mods[i], nn.ReLU):
                relu = True
                i += 1
            O, I, _ = w.shape
            ipad = _round_up(I) if in_pad is None else in_pad       # the previous conv's padded output feeds this one
            opad = _round_up(O)
            w4 = torch.zeros((opad, ipad, 1, 3), device=device)
            w4[:O, :I, 0, :] = w
            b4 = torch.zeros(opad, device=device)
            b4[:O] = b
            self.convs.append((w4.contiguous(memory_format=torch.channels_last), b4, relu, O))
            in_pad = opad
        self.in_channels = self.convs[0][0].shape[1]                # padded
        self.out_channels = self.convs[-1][3]

    def __call__(self, x, fused_ok):
        """x: [B, P, I_pad] (NHWC storage) -> [B, P, O_pad]"""
        B, P, C = x.shape
        h = x.view(B, 1, P, C).permute(0, 3, 1, 2)                  # logical NCHW, channels-last strides
        for w, b, relu, _ in self.convs:
            if relu and fused_ok:
                h = torch.cudnn_convolution_relu(h, w, b, (1, 1), (0, 1), (1, 1), 1)
            else:
                h = F.conv2d(h, w, b, padding=(0, 1))
                if relu:
                    h = torch.relu_(h)
        return h.permute(0, 2, 3, 1).reshape(B, P, h.shape[1])


class FusedDecoder:
    """``FusedDecoder(network.decoder)(z [B, latent]) -> [B, M, 3]`` (see the module docstring)."""

    def __init__(self, decoder, device=None):
        if decoder.training:
            raise RuntimeError("FusedDecoder folds eval-mode BatchNorm statistics: call network.eval() first")
        p = next(decoder.parameters())
        self.device = torch.device(device) if device is not None else p.device
        self.latent = decoder.latent_dimension
        self.layers = []
        for name in ("layer1", "layer2", "layer3", "layer4"):
            L = getattr(decoder, name)
            grid = L.grid.detach().to(self.device, torch.float32).reshape(1, -1, 1)      # [1, P_out, 1]
            self.layers.append((L.m, grid, _Fold(L.fold1, self.device), _Fold(L.fold2, self.device)))
        self.fused_ok = self.device.type == "cuda" and hasattr(torch, "cudnn_convolution_relu")

    @staticmethod
    def _assemble(parts, width):
        """Concatenate [B, P, c_k] pieces along the channel axis and zero-pad to `width` channels."""
        B, P = parts[0].shape[:2]
        have = sum(t.shape[2] for t in parts)
        if have < width:
            parts = parts + [parts[0].new_zeros(1, 1, 1).expand(B, P, width - have)]
        return torch.cat(parts, dim=2)

    @torch.no_grad()
    def __call__(self, z):
        z = z.reshape(-1, self.latent).to(self.device, torch.float32)
        B = z.shape[0]
        h = z.view(B, 1, self.latent)                                # [B, P = 1, C]
        for m, grid, fold1, fold2 in self.layers:
            P = h.shape[1] * m
            x = h.unsqueeze(2).expand(-1, -1, m, -1).reshape(B, P, h.shape[2])      # repeat_interleave along the points
            g = grid.expand(B, -1, -1)
            try:
                first = fold1(self._assemble([g, x], fold1.in_channels), self.fused_ok)[:, :, :fold1.out_channels]
                second = fold2(self._assemble([g, x, first], fold2.in_channels), self.fused_ok)[:, :, :fold2.out_channels]
            except torch.cuda.OutOfMemoryError:
                raise                                                # not a missing kernel: the caller must see it
            except RuntimeError:
                if not self.fused_ok:
                    raise
                self.fused_ok = False                                # this cuDNN build has no fused conv+bias+relu for the shape
                first = fold1(self._assemble([g, x], fold1.in_channels), False)[:, :, :fold1.out_channels]
                second = fold2(self._assemble([g, x, first], fold2.in_channels), False)[:, :, :fold2.out_channels]
            h = first + second
        return h                                                     # [B, M, 3]


def fuse_decoder(network):
    """A ``decode(z) -> [B, M, 3]`` callable for a FoldingNet-style autoencoder (``network.decoder``), or ``None``
    if the network does not have that structure (the caller then keeps ``network.decode``)."""
    dec = getattr(network, "decoder", None)
    try:
        if dec is None or next(dec.parameters()).dtype != torch.float32:     # the fused path computes in fp32 / TF32 only
            return None
        return FusedDecoder(dec)
    except (AttributeError, TypeError, StopIteration):
        return None

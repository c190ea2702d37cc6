"""FoldingNet-style autoencoder used as the decoder of the latent-grid scan and the training step.

The network is NOT a kernel target of this project (its Conv1d/BatchNorm stack already runs on
cuDNN/cuBLAS): it exists here so that BASELINE configs 2, 3 and 5 can
run on a machine that has no molearn installation.  The architecture and the parameter names
follow the reference (src/molearn/models/foldingnet.py:52-255) so that ``state_dict``s are
interchangeable in both directions - ``tests/test_models.py`` loads one into the other and
compares outputs when a reference tree is present.

Interface (foldingnet.py:235-255): ``encode([B, N, 3]) -> [B, latent]``,
``decode([B, latent]) -> [B, M, 3]`` with ``M = (N // 128 + 1) * 128 >= N``; the decoder output
is a permuted view of a ``[B, 3, M]`` tensor, i.e. it has strides ``(3M, 1, M)``.
"""
from __future__ import annotations

import torch
from torch import nn


def _neighbours(feat, k):
    """Indices [B, N, k] of the k nearest points (squared euclidean, self included). feat: [B, C, N]."""
    sq = (feat * feat).sum(dim=1)                                    # [B, N]
    gram = torch.bmm(feat.transpose(1, 2), feat)                      # [B, N, N]
    neg_d2 = 2.0 * gram - sq.unsqueeze(2) - sq.unsqueeze(1)
    return neg_d2.topk(k, dim=-1).indices


def _gather_points(points, idx):
    """points [B, N, C], idx [B, N, k] -> [B, N, k, C]."""
    B, N, k = idx.shape
    flat = idx.reshape(B, N * k, 1).expand(-1, -1, points.shape[-1])
    return points.gather(1, flat).view(B, N, k, points.shape[-1])


class GraphLayer(nn.Module):
    """max-pool over the kNN graph, then a pointwise conv + BN + ReLU (foldingnet.py:33-50)."""

    def __init__(self, in_channel, out_channel, k=16):
        super().__init__()
        self.k = k
        self.conv = nn.Conv1d(in_channel, out_channel, 1)
        self.bn = nn.BatchNorm1d(out_channel)

    def forward(self, x):
        idx = _neighbours(x, self.k)
        pooled = _gather_points(x.transpose(1, 2), idx).amax(dim=2).transpose(1, 2)
        return torch.relu(self.bn(self.conv(pooled)))


class Encoder(nn.Module):
    """foldingnet.py:52-100."""

    def __init__(self, latent_dimension=2, **_):
        super().__init__()
        self.latent_dimension = latent_dimension
        widths = [12, 64, 64, 64]
        for i in range(3):
            setattr(self, f"conv{i + 1}", nn.Conv1d(widths[i], widths[i + 1], 1))
            setattr(self, f"bn{i + 1}", nn.BatchNorm1d(widths[i + 1]))
        self.graph_layer1 = GraphLayer(64, 128, k=16)
        self.graph_layer2 = GraphLayer(128, 1024, k=16)
        self.conv4 = nn.Conv1d(1024, 512, 1)
        self.bn4 = nn.BatchNorm1d(512)
        self.conv5 = nn.Conv1d(512, latent_dimension, 1)

    def forward(self, x):                                             # x [B, 3, N]
        B, _, N = x.shape
        local = _gather_points(x.transpose(1, 2), _neighbours(x, 16))  # [B, N, 16, 3]
        local = local - local.mean(dim=2, keepdim=True)
        cov = torch.matmul(local.transpose(2, 3), local).reshape(B, N, 9).transpose(1, 2)
        h = torch.cat([x, cov], dim=1)
        for i in (1, 2, 3):
            h = torch.relu(getattr(self, f"bn{i}")(getattr(self, f"conv{i}")(h)))
        h = self.graph_layer2(self.graph_layer1(h))
        h = self.bn4(self.conv4(h)).amax(dim=-1, keepdim=True)
        return self.conv5(h).squeeze(-1)


class FoldingLayer(nn.Module):
    """Conv1d(k=3)+BN+ReLU stack ending in a plain Conv1d (foldingnet.py:103-127)."""

    def __init__(self, in_channel, out_channels):
        super().__init__()
        mods, c = [], in_channel
        for oc in out_channels[:-1]:
            mods += [nn.Conv1d(c, oc, 3, 1, 1), nn.BatchNorm1d(oc), nn.ReLU(inplace=True)]
            c = oc
        mods.append(nn.Conv1d(c, out_channels[-1], 3, 1, 1))
        self.layers = nn.Sequential(*mods)

    def forward(self, *parts):
        return self.layers(torch.cat(parts, dim=1))


class Decoder_Layer(nn.Module):
    """One up-sampling stage: repeat every point m times, fold twice over a 1-D grid (foldingnet.py:130-159)."""

    def __init__(self, in_points, out_points, in_channel, out_channel, **_):
        super().__init__()
        assert out_points % in_points == 0
        self.out_points = out_points
        self.m = out_points // in_points
        self.grid = torch.linspace(-0.5, 0.5, out_points).view(1, -1)  # plain attribute, like the reference
        self.fold1 = FoldingLayer(in_channel + 1, [512, 512, out_channel])
        self.fold2 = FoldingLayer(in_channel + out_channel + 1, [512, 512, out_channel])

    def forward(self, x):
        if self.grid.device != x.device:
            self.grid = self.grid.to(x.device)
        grid = self.grid.unsqueeze(0).expand(x.shape[0], -1, -1)
        x = x.repeat_interleave(self.m, dim=-1)
        first = self.fold1(grid, x)
        return first + self.fold2(grid, x, first)


class Decoder(nn.Module):
    """foldingnet.py:162-189: 1 -> s -> 8s -> 32s -> 128s points, s = out_points // 128 + 1."""

    def __init__(self, out_points, latent_dimension=2, **_):
        super().__init__()
        self.latent_dimension = latent_dimension
        self.out_points = out_points
        s = out_points // 128 + 1
        self.layer1 = Decoder_Layer(1, s, latent_dimension, 3 * 128)
        self.layer2 = Decoder_Layer(s, s * 8, 3 * 128, 3 * 16)
        self.layer3 = Decoder_Layer(s * 8, s * 32, 3 * 16, 3 * 4)
        self.layer4 = Decoder_Layer(s * 32, s * 128, 3 * 4, 3)

    def forward(self, z):
        h = z.reshape(-1, self.latent_dimension, 1)
        return self.layer4(self.layer3(self.layer2(self.layer1(h))))


class AutoEncoder(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.encoder = Encoder(*args, **kwargs)
        self.decoder = Decoder(*args, **kwargs)

    def encode(self, x):                                              # [B, N, 3] -> [B, latent]
        return self.encoder(x.permute(0, 2, 1))

    def decode(self, z):                                              # [B, latent] -> [B, M, 3]
        return self.decoder(z).permute(0, 2, 1)

    def forward(self, x):
        return self.decode(self.encode(x))

"""Layers of the EDM DDPM++ network on the xdb200 kernels (reference: layers/edm.py:59-343).

``Linear`` / ``Conv2d`` / ``GroupNorm`` are parameter holders with the reference's names, shapes and buffers (the
``resample_filter`` buffer of resampling convolutions is part of the reference's state dict); all arithmetic happens in
``UNetBlock.forward`` on NHWC bf16 activations:

    GN32 + SiLU -> (nearest x2 | 2x2 average) -> conv3x3 (+ bias + affine(emb): the noise embedding is ONE row for the whole
    batch, so the per-channel term is folded into the conv bias) -> GN32 + SiLU -> conv3x3 with the skip path as an extra K
    segment of the same tensor-core tile (1x1 projection, or sqrt(1/2) * I where the reference adds the input unchanged) ->
    [GN32 -> qkv 1x1 -> one 256-wide attention head -> proj 1x1 with the residual as a second K segment].

The ``skip_scale`` of the reference (sqrt(1/2) after every residual sum) is folded into the packed weights and biases.
"""
import math

import torch

from .. import ops
from .resnet import _as_samples
from .utils import Packed, pack_conv3x3


def _xavier_uniform(shape, fan_in, fan_out):
    return math.sqrt(6 / (fan_in + fan_out)) * (torch.rand(shape) * 2 - 1)


class Linear(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=True, init_weight=1.0, **_):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = torch.nn.Parameter(_xavier_uniform([out_features, in_features], in_features, out_features) * init_weight)
        self.bias = torch.nn.Parameter(torch.zeros(out_features)) if bias else None


class Conv2d(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel, bias=True, up=False, down=False, resample_filter=(1, 1),
                 fused_resample=False, init_weight=1.0, **_):
        super().__init__()
        if fused_resample or list(resample_filter) != [1, 1] or kernel not in (1, 3):
            raise NotImplementedError("only the DDPM++ configuration (box resampling filter, kernel 1 / 3)")
        self.in_channels, self.out_channels, self.up, self.down = in_channels, out_channels, up, down
        fan = kernel * kernel
        self.weight = torch.nn.Parameter(
            _xavier_uniform([out_channels, in_channels, kernel, kernel], in_channels * fan, out_channels * fan) * init_weight)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        f = torch.as_tensor(list(resample_filter), dtype=torch.float32)
        self.register_buffer("resample_filter", f.ger(f)[None, None] / f.sum().square() if up or down else None)


class GroupNorm(torch.nn.Module):
    def __init__(self, num_channels, num_groups=32, min_channels_per_group=4, eps=1e-5):
        super().__init__()
        self.num_groups, self.eps = min(num_groups, num_channels // min_channels_per_group), eps
        if self.num_groups != 32:
            raise NotImplementedError("GroupNorm with other than 32 groups")
        self.weight = torch.nn.Parameter(torch.ones(num_channels))
        self.bias = torch.nn.Parameter(torch.zeros(num_channels))


def _resample(x, up, down):
    if not (up or down):
        return x
    n, H, W, C = x.shape
    out = torch.empty((n, H * 2, W * 2, C) if up else (n, H // 2, W // 2, C), device=x.device, dtype=torch.bfloat16)
    (torch.ops.xdb200.upsample2x if up else torch.ops.xdb200.avgpool2x2)(x, out)
    return out


def _rows(x):
    """NHWC view with dense pixels -> [pixels, C] view (row stride = pixel stride)."""
    n, H, W, C = x.shape
    return x.as_strided((n * H * W, C), (x.stride(2), 1))


class UNetBlock(torch.nn.Module, Packed):
    def __init__(self, in_channels, out_channels, emb_channels, up=False, down=False, attention=False, num_heads=None,
                 channels_per_head=64, dropout=0, skip_scale=1, eps=1e-5, resample_filter=(1, 1), resample_proj=False,
                 adaptive_scale=True, init=None, init_zero=None, init_attn=None):
        super().__init__()
        if adaptive_scale:
            raise NotImplementedError("adaptive_scale=True (ADM blocks)")
        self.in_channels, self.out_channels, self.up, self.down = in_channels, out_channels, up, down
        self.num_heads = 0 if not attention else (num_heads if num_heads is not None else out_channels // channels_per_head)
        if self.num_heads not in (0, 1) or (self.num_heads and out_channels != 256):
            raise NotImplementedError("attention other than one 256-wide head")
        self.skip_scale = float(skip_scale)
        self.norm0 = GroupNorm(in_channels, eps=eps)
        self.conv0 = Conv2d(in_channels, out_channels, 3, up=up, down=down, resample_filter=resample_filter)
        self.affine = Linear(emb_channels, out_channels)
        self.norm1 = GroupNorm(out_channels, eps=eps)
        self.conv1 = Conv2d(out_channels, out_channels, 3, init_weight=1e-5)
        self.skip = None
        if out_channels != in_channels or up or down:
            if not (resample_proj or out_channels != in_channels):
                raise NotImplementedError("resampling skip without a projection")
            self.skip = Conv2d(in_channels, out_channels, 1, up=up, down=down, resample_filter=resample_filter)
        if self.num_heads:
            self.norm2 = GroupNorm(out_channels, eps=eps)
            self.qkv = Conv2d(out_channels, out_channels * 3, 1, init_weight=math.sqrt(0.2))
            self.proj = Conv2d(out_channels, out_channels, 1, init_weight=1e-5)

    def _packs(self):
        s, co = self.skip_scale, self.out_channels
        params = [self.conv0.weight, self.conv1.weight, self.conv1.bias]
        if self.skip is not None:
            params += [self.skip.weight, self.skip.bias]
        if self.num_heads:
            params += [self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias]

        def build():
            dev = self.conv0.weight.device
            eye = torch.eye(co, device=dev)
            w0 = pack_conv3x3(self.conv0.weight)
            wskip = self.skip.weight.reshape(co, -1) if self.skip is not None else eye
            w1 = pack_conv3x3(self.conv1.weight * s, wskip * s)
            b1 = ((self.conv1.bias + (self.skip.bias if self.skip is not None else 0)) * s).float().contiguous()
            out = {"w0": w0, "w1": w1, "b1": b1}
            if self.num_heads:
                # the reference splits the 3C output channels as (c, j): channel 3c + j is q / k / v [j] of head channel c;
                # rows are re-ordered to [Q | K | V] so that each operand is one contiguous 256-wide slab per token
                wq = self.qkv.weight.reshape(co, 3, co).permute(1, 0, 2).reshape(3 * co, co)
                out["wqkv"] = wq.to(torch.bfloat16).contiguous()
                out["bqkv"] = self.qkv.bias.reshape(co, 3).t().reshape(-1).float().contiguous()
                out["wproj"] = (torch.cat([self.proj.weight.reshape(co, co), eye], 1) * s).to(torch.bfloat16).contiguous()
                out["bproj"] = (self.proj.bias * s).float().contiguous()
            return out
        return self.packed("w", tuple(params), build)

    def forward(self, x, emb_bias, samples, out=None):
        """x bf16 NHWC view [n, H, W, Cin]; emb_bias fp32 [Cout] = conv0.bias + affine(emb) (one row for the batch)."""
        pk = self._packs()
        h = ops.groupnorm(_as_samples(x, samples), self.norm0.weight, self.norm0.bias, eps=self.norm0.eps, silu=True).view(x.shape)
        h = ops.conv3x3(_resample(h, self.up, self.down), pk["w0"], emb_bias, qstats=True)
        h = ops.groupnorm(_as_samples(h, samples), self.norm1.weight, self.norm1.bias, eps=self.norm1.eps, silu=True).view(h.shape)
        y = ops.conv3x3(h, pk["w1"], pk["b1"], xs=_resample(x, self.up, self.down), out=None if self.num_heads else out,
                        qstats=True)
        if not self.num_heads:
            return y
        n, H, W, C = y.shape
        hn = ops.groupnorm(_as_samples(y, samples), self.norm2.weight, self.norm2.bias, eps=self.norm2.eps).view(y.shape)
        qkv = ops.linear(_rows(hn), pk["wqkv"], pk["bqkv"])
        v5 = qkv.view(n, H * W, 3, 1, C)
        q, k, v = (v5[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        a = ops.attention(q, k, v, 1 / math.sqrt(C))                               # [n, 1, T, C] view of [n, T, 1, C]
        if out is None:
            out = torch.empty_like(y)
        ops.linear(a.permute(0, 2, 1, 3).reshape(n * H * W, C), pk["wproj"], pk["bproj"], a2=_rows(y), out=_rows(out), qstats=True)
        return out


"""BigGAN-style residual block, Downsample, Upsample on the xdb200 kernels (NHWC bf16).
Parameter names / constructor arguments follow the reference (layers/resnet.py:83-201,440-502)."""
from typing import Dict

import torch

from .. import ops
from .utils import ContextBlock, Packed, pack_conv3x3, zero_module


def _as_samples(x, samples):
    """NHWC view [nimg,H,W,C] -> [samples, pixels, C] (GroupNorm statistics unit), stride-preserving."""
    nimg, H, W, C = x.shape
    return x.as_strided((samples, nimg // samples * H * W, C),
                        (x.stride(0) * (nimg // samples), x.stride(2), 1))


class ResnetBlockBigGAN(ContextBlock, Packed):
    """GN32+SiLU -> conv3x3 -> GN32*(1+scale)+shift -> SiLU -> conv3x3 (zero-init) -> + skip(x).
    The time-embedding linear of ALL blocks is batched by the owning network (same input for every
    block); this block receives its [scale | shift] slice.  The 1x1 skip projection is accumulated
    inside the second conv's tensor-core tile (extra K segment) instead of a separate GEMM + add."""
    conv_dims = 2

    def __init__(self, dim_in, time_emb_dim, dropout, dim_out=None, use_conv=False, use_scale_shift_norm=False,
                 up=False, down=False, **kwargs):
        super().__init__()
        if up or down:
            raise NotImplementedError("resblock_updown=True")
        if not use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=False")
        if use_conv:
            raise NotImplementedError("3x3 skip connection (resamp_with_conv=True)")
        self.input_channels, self.emb_channels = dim_in, time_emb_dim
        self.out_channels = dim_out or dim_in
        conv = torch.nn.Conv2d if self.conv_dims == 2 else torch.nn.Conv3d
        k, pad = (3, 1) if self.conv_dims == 2 else ((1, 3, 3), (0, 1, 1))
        self.in_layers = torch.nn.Sequential(torch.nn.GroupNorm(32, dim_in), torch.nn.SiLU(),
                                             conv(dim_in, self.out_channels, k, padding=pad))
        self.emb_layers = self._make_emb_layers(time_emb_dim, 2 * self.out_channels, kwargs)
        self.out_layers = torch.nn.Sequential(
            torch.nn.GroupNorm(32, self.out_channels), torch.nn.SiLU(), torch.nn.Dropout(p=dropout),
            zero_module(conv(self.out_channels, self.out_channels, k, padding=pad)))
        self.skip_connection = torch.nn.Identity() if self.out_channels == dim_in \
            else conv(dim_in, self.out_channels, 1)

    @staticmethod
    def _make_emb_layers(time_emb_dim, width, kwargs):
        return torch.nn.Sequential(torch.nn.SiLU(), torch.nn.Linear(time_emb_dim, width))

    def emb_linear(self):
        """(weight, bias) of the final time-embedding linear, for the network-level batched GEMM."""
        return self.emb_layers[1].weight, self.emb_layers[1].bias

    def forward(self, x, scale_shift, samples, out=None):
        """x bf16 NHWC [nimg,H,W,Cin] (may be a concat buffer); scale_shift fp32 [samples, 2*Cout] view;
        ``samples`` = number of GroupNorm statistics units (B; frames of a clip share statistics)."""
        c1, c2, skip = self.in_layers[2], self.out_layers[3], self.skip_connection
        has_skip = not isinstance(skip, torch.nn.Identity)
        params = (c1.weight, c2.weight) + ((skip.weight, skip.bias, c2.bias) if has_skip else ())
        w1, w2, b2 = self.packed("w", params, lambda: (
            pack_conv3x3(c1.weight), pack_conv3x3(c2.weight, skip.weight if has_skip else None),
            (c2.bias + skip.bias).detach().float() if has_skip else None))
        g1, g2 = self.in_layers[0], self.out_layers[0]
        h = ops.groupnorm(_as_samples(x, samples), g1.weight, g1.bias, eps=g1.eps, silu=True).view(x.shape)
        h = ops.conv3x3(h, w1, c1.bias)
        h = ops.groupnorm(_as_samples(h, samples), g2.weight, g2.bias, scale_shift=scale_shift, eps=g2.eps,
                          silu=True).view(h.shape)
        if has_skip:
            return ops.conv3x3(h, w2, b2, xs=x, out=out)
        return ops.conv3x3(h, w2, c2.bias, residual=x, out=out)


class Downsample(torch.nn.Module):
    """AvgPool 2x2 over (H, W) (reference: resnet.py:440-467; dims=3 pools (1,2,2))."""

    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        if use_conv:
            raise NotImplementedError("resamp_with_conv=True")
        self.channels, self.dims = channels, dims

    def forward(self, x, out=None):
        nimg, H, W, C = x.shape
        if out is None:
            out = torch.empty((nimg, H // 2, W // 2, C), device=x.device, dtype=torch.bfloat16)
        torch.ops.xdb200.avgpool2x2(x, out)
        return out


class Upsample(torch.nn.Module):
    """Nearest x2 over (H, W) (reference: resnet.py:470-502)."""

    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        if use_conv:
            raise NotImplementedError("resamp_with_conv=True")
        self.channels, self.dims = channels, dims

    def forward(self, x, out=None):
        nimg, H, W, C = x.shape
        if out is None:
            out = torch.empty((nimg, H * 2, W * 2, C), device=x.device, dtype=torch.bfloat16)
        torch.ops.xdb200.upsample2x(x, out)
        return out

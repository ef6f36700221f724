"""BigGAN-style residual block, Downsample, Upsample on the xdb200 kernels (NHWC bf16).
Parameter names / constructor arguments follow the reference (layers/resnet.py:83-201,440-502)."""
import os
from typing import Dict

import torch

from .. import ops
from .utils import ContextBlock, Packed, pack_conv3x3, zero_module


# First conv of a resblock + its GroupNorm as one library call (xd_conv3x3_groupnorm_bf16_tc: at the 8x8 / 4x4 levels the
# split-K reduce pass normalises).  Measured neutral on the UNet step (37.4 vs 36.9 img/s: one CTA per sample sums the partial
# tiles more slowly than the wide reduce kernel + the 5 us cluster GroupNorm it replaces), so it is an opt-in.
CONV_GN_FUSED = os.environ.get("XDB200_CONV_GN", "0") == "1"


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
        """x bf16 NHWC [nimg,H,W,Cin] (may be a concat buffer); scale_shift fp32 [samples, 2*Cout] view (or a callable
        returning it, resolved after the first convolution has been launched);
        ``samples`` = number of GroupNorm statistics units (B; frames of a clip share statistics)."""
        c1, c2, skip = self.in_layers[2], self.out_layers[3], self.skip_connection
        has_skip = not isinstance(skip, torch.nn.Identity)
        params = (c1.weight, c2.weight) + ((skip.weight, skip.bias, c2.bias) if has_skip else ())
        w1, w2, b2 = self.packed("w", params, lambda: (
            pack_conv3x3(c1.weight), pack_conv3x3(c2.weight, skip.weight if has_skip else None),
            (c2.bias + skip.bias).detach().float() if has_skip else None))
        g1, g2 = self.in_layers[0], self.out_layers[0]
        h = ops.groupnorm(_as_samples(x, samples), g1.weight, g1.bias, eps=g1.eps, silu=True).view(x.shape)
        late = callable(scale_shift)             # conditioning branch on a side stream: join as late as possible
        if late and (ops.MATMUL_BACKEND == "tc" and CONV_GN_FUSED and self.out_channels % 128 == 0):
            scale_shift, late = scale_shift(), False
        if ops.MATMUL_BACKEND == "tc" and CONV_GN_FUSED and self.out_channels % 128 == 0:
            # first conv + the GroupNorm that consumes it as one library call (split-K reduce that normalises at the low
            # resolutions, statistics from the conv epilogue at the high ones: include/xdb200.h)
            h = ops.conv3x3_groupnorm(h, w1, c1.bias, samples, g2.weight, g2.bias, scale_shift=scale_shift, eps=g2.eps)
        else:
            h = ops.conv3x3(h, w1, c1.bias, qstats=True)
            if late:
                scale_shift = scale_shift()
            h = ops.groupnorm(_as_samples(h, samples), g2.weight, g2.bias, scale_shift=scale_shift, eps=g2.eps,
                              silu=True).view(h.shape)
        if has_skip:
            return ops.conv3x3(h, w2, b2, xs=x, out=out, qstats=True)
        return ops.conv3x3(h, w2, c2.bias, residual=x, out=out, qstats=True)


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


class Upsample(torch.nn.Module, Packed):
    """Nearest x2 over (H, W), optionally followed by a 3x3 convolution (reference: resnet.py:470-502)."""

    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        if use_conv and dims != 2:
            raise NotImplementedError("resamp_with_conv=True for video")
        self.channels, self.dims, self.use_conv = channels, dims, use_conv
        if use_conv:
            self.conv = torch.nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, out=None):
        nimg, H, W, C = x.shape
        up = out if out is not None and not self.use_conv else \
            torch.empty((nimg, H * 2, W * 2, C), device=x.device, dtype=torch.bfloat16)
        ops.upsample2x(x, up)
        if not self.use_conv:
            return up
        w = self.packed("w", (self.conv.weight,), lambda: pack_conv3x3(self.conv.weight))
        return ops.conv3x3(up, w, self.conv.bias, out=out)


class ResnetBlockEfficient(ContextBlock, Packed):
    """Imagen "efficient" residual block (reference: resnet.py:204-250, figure A.27): GN32 + SiLU -> conv3x3 -> GN32 + SiLU ->
    conv3x3 (zero-init), always a 1x1 skip projection, (skip + h) * 0.7071, no time embedding.  The skip projection is an
    extra K segment of the second conv's tensor-core tile and the 0.7071 is folded into its weights and bias."""
    SKIP_SCALE = 0.7071

    def __init__(self, dim_in, dropout=0.0, dim_out=None, scale_skip_connection: bool = True, **kwargs):
        super().__init__()
        self._input_channels, self._output_channels = dim_in, dim_out or dim_in
        co = self._output_channels
        self._resnet_path = torch.nn.Sequential(
            torch.nn.GroupNorm(32, dim_in), torch.nn.SiLU(), torch.nn.Conv2d(dim_in, co, 3, padding=1),
            torch.nn.GroupNorm(32, co), torch.nn.SiLU(), torch.nn.Dropout(p=dropout),
            zero_module(torch.nn.Conv2d(co, co, 3, padding=1)))
        self._scale_skip_connection = scale_skip_connection
        self._skip_connection = torch.nn.Conv2d(dim_in, co, 1)

    def forward(self, x, samples, out=None):
        g1, c1, g2, c2, skip = self._resnet_path[0], self._resnet_path[2], self._resnet_path[3], self._resnet_path[6], \
            self._skip_connection
        s = self.SKIP_SCALE if self._scale_skip_connection else 1.0
        w1, w2, b2 = self.packed("w", (c1.weight, c2.weight, c2.bias, skip.weight, skip.bias), lambda: (
            pack_conv3x3(c1.weight), pack_conv3x3(c2.weight * s, skip.weight * s), ((c2.bias + skip.bias) * s).detach().float()))
        h = ops.groupnorm(_as_samples(x, samples), g1.weight, g1.bias, eps=g1.eps, silu=True).view(x.shape)
        h = ops.conv3x3(h, w1, c1.bias, qstats=True)
        h = ops.groupnorm(_as_samples(h, samples), g2.weight, g2.bias, eps=g2.eps, silu=True).view(h.shape)
        return ops.conv3x3(h, w2, b2, xs=x, out=out, qstats=True)


class _EfficientStage(ContextBlock, Packed):
    """Shared body of DBlock / UBlock: + Linear(SiLU(temb)) per channel -> ResnetBlockEfficient chain -> optional attention."""

    def _build(self, dim_in, dim_out, num_resnet_blocks, time_embedding_dim, attention_type, attention_kwargs, dropout):
        self._input_channels, self._output_channels = dim_in, dim_out or dim_in
        self._embedding_layers = torch.nn.Sequential(torch.nn.SiLU(), torch.nn.Linear(time_embedding_dim, dim_in))
        self._resnet_blocks = torch.nn.Sequential(*[
            ResnetBlockEfficient(dim_in=dim_in if i == 0 else self._output_channels, dim_out=self._output_channels,
                                 dropout=dropout) for i in range(num_resnet_blocks)])
        self._attention = attention_type(in_channels=self._output_channels, **attention_kwargs) \
            if attention_type is not None else torch.nn.Identity()

    def emb_linear(self):
        return self._embedding_layers[1].weight, self._embedding_layers[1].bias

    def _body(self, h, emb, samples, context):
        """h bf16 NHWC [n, H, W, C_in]; emb fp32 [samples, C_in] = this stage's slice of the batched embedding GEMM."""
        n, H, W, C = h.shape
        biased = torch.empty((n, H, W, C), device=h.device, dtype=torch.bfloat16)
        torch.ops.xdb200.add_channel_bias(_as_samples(h, samples), emb, biased.view(samples, -1, C))
        h = biased
        for rb in self._resnet_blocks:
            h = rb(h, samples)
        if not isinstance(self._attention, torch.nn.Identity):
            h = self._attention(h, context=context)
        return h


class DBlock(_EfficientStage):
    """reference: resnet.py:253-330 (figure A.28): stride-2 conv3x3 first (im2col + tensor-core GEMM)."""

    def __init__(self, dim_in, num_resnet_blocks, time_embedding_dim, downsample: bool, attention_type=None,
                 attention_kwargs=None, dropout=0.0, dim_out=None, **kwargs):
        super().__init__()
        self._downsampling_convolution = torch.nn.Conv2d(dim_in, dim_in, 3, padding=1, stride=2) if downsample \
            else torch.nn.Identity()
        self._build(dim_in, dim_out, num_resnet_blocks, time_embedding_dim, attention_type, attention_kwargs or {}, dropout)

    def forward(self, x, emb, samples, context):
        conv = self._downsampling_convolution
        if not isinstance(conv, torch.nn.Identity):
            n, H, W, C = x.shape
            w = self.packed("wd", (conv.weight,), lambda: pack_conv3x3(conv.weight))
            cols = torch.empty((n * (H // 2) * (W // 2), 9 * C), device=x.device, dtype=torch.bfloat16)
            torch.ops.xdb200.im2col3x3_s2(x, cols)
            x = ops.linear(cols, w, conv.bias).view(n, H // 2, W // 2, C)
        return self._body(x, emb, samples, context)


class UBlock(_EfficientStage):
    """reference: resnet.py:333-437 (figure A.29): nearest x2 + conv3x3 last."""

    def __init__(self, dim_in, num_resnet_blocks, time_embedding_dim, upsample: bool, attention_type=None,
                 attention_kwargs=None, dropout=0.0, dim_out=None, **kwargs):
        super().__init__()
        self._upsample = Upsample(channels=dim_out or dim_in, use_conv=True) if upsample else torch.nn.Identity()
        self._build(dim_in, dim_out, num_resnet_blocks, time_embedding_dim, attention_type, attention_kwargs or {}, dropout)

    def forward(self, x, emb, samples, context):
        h = self._body(x, emb, samples, context)
        return h if isinstance(self._upsample, torch.nn.Identity) else self._upsample(h)

"""Attention blocks on the xdb200 kernels.  Parameter names / constructor arguments follow the
reference (xdiffusion/layers/attention.py) for YAML + checkpoint compatibility."""
import math
from typing import Dict, Optional

import torch

from .. import _lib, ops
from .utils import ContextBlock, Packed, bf16_weight, zero_module


class MultiHeadSelfAttention(torch.nn.Module, Packed):
    """DiT / PixArt self-attention over token rows: qkv [Q|K|V]-major, softmax(q k^T / sqrt(d)) v, proj
    (reference: attention.py:313-380, unfused branch)."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False, **kwargs):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = torch.nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = torch.nn.Linear(dim, dim)

    def weights(self):
        return self.packed("w", (self.qkv.weight, self.proj.weight),
                           lambda: (bf16_weight(self.qkv.weight), bf16_weight(self.proj.weight)))

    def head_packed(self):
        """qkv Linear re-packed per head as [q_h | k_h | v_h] rows (bf16 [H*3*d, D]) + bias (fp32 [H*3*d]): one head's
        projections are then one contiguous 192-row weight tile (csrc/dit_block.cu)."""
        def build():
            H, d = self.num_heads, self.head_dim
            D = H * d
            w = self.qkv.weight.detach().view(3, H, d, D).permute(1, 0, 2, 3).reshape(3 * D, D)
            b = self.qkv.bias.detach().float().view(3, H, d).permute(1, 0, 2).reshape(3 * D)
            return w.to(torch.bfloat16).contiguous(), b.contiguous()
        return self.packed("wh", (self.qkv.weight, self.qkv.bias), build)

    def attend(self, x, tokens: int, ln=None):
        """softmax(q k^T / sqrt(d)) v for every head, WITHOUT the output projection: bf16 [B*T, D]."""
        wq, _ = self.weights()
        M, D = x.shape
        B, H, d = M // tokens, self.num_heads, self.head_dim
        qkv = ops.ln_linear(x, ln[0], ln[1], ln[2], wq, self.qkv.bias) if ln is not None else ops.linear(x, wq, self.qkv.bias)
        qkv = qkv.view(B, tokens, 3, H, d)
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        return ops.attention(q, k, v, self.scale).permute(0, 2, 1, 3).reshape(M, D)

    def forward(self, x_bf16, tokens: int, ln=None, **epilogue):
        """x bf16 [B*T, D]; ``epilogue`` (gate / residual / out) is fused into the proj GEMM.
        ``ln=(shift, scale, rows_per_mod)``: x is the fp32 residual stream, LayerNorm + modulate is fused into qkv."""
        wq, wp = self.weights()
        M, D = x_bf16.shape
        B, H, d = M // tokens, self.num_heads, self.head_dim
        if ln is not None:
            qkv = ops.ln_linear(x_bf16, ln[0], ln[1], ln[2], wq, self.qkv.bias)
        else:
            qkv = ops.linear(x_bf16, wq, self.qkv.bias)
        qkv = qkv.view(B, tokens, 3, H, d)
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        a = ops.attention(q, k, v, self.scale)                       # [B,H,T,d] view of [B,T,H*d]
        a2 = a.permute(0, 2, 1, 3).reshape(M, D)
        return ops.linear(a2, wp, self.proj.bias, **epilogue)


class LastChannelCrossAttention(torch.nn.Module, Packed):
    """PixArt cross-attention: q from x, k/v from the context tokens, no mask, no q/k/v bias
    (reference: attention.py:191-228)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        context_dim = context_dim if context_dim is not None else query_dim
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        self.to_q = torch.nn.Linear(query_dim, inner, bias=False)
        self.to_k = torch.nn.Linear(context_dim, inner, bias=False)
        self.to_v = torch.nn.Linear(context_dim, inner, bias=False)
        self.to_out = torch.nn.Linear(inner, query_dim)

    def project_context(self, y_bf16):
        """k|v for the context tokens: y bf16 [B, L, C] -> bf16 [B, L, 2, H, d].  Timestep-invariant."""
        wkv = self.packed("wkv", (self.to_k.weight, self.to_v.weight),
                          lambda: torch.cat([bf16_weight(self.to_k.weight), bf16_weight(self.to_v.weight)], 0))
        B, L, C = y_bf16.shape
        return ops.linear(y_bf16.reshape(B * L, C), wkv).view(B, L, 2, self.heads, self.dim_head)

    def weights(self):
        """(to_q, to_out) as bf16 [N, K] GEMM operands."""
        return self.packed("w", (self.to_q.weight, self.to_out.weight),
                           lambda: (bf16_weight(self.to_q.weight), bf16_weight(self.to_out.weight)))

    def attend(self, x_bf16, tokens: int, kv):
        """softmax(q k^T / sqrt(d)) v WITHOUT the output projection: bf16 [B*T, H*d]."""
        wq, _ = self.weights()
        M, D = x_bf16.shape
        B, H, d = M // tokens, self.heads, self.dim_head
        q = ops.linear(x_bf16, wq).view(B, tokens, H, d).permute(0, 2, 1, 3)
        k, v = kv[:, :, 0].permute(0, 2, 1, 3), kv[:, :, 1].permute(0, 2, 1, 3)
        return ops.attention(q, k, v, self.scale).permute(0, 2, 1, 3).reshape(M, H * d)

    def forward(self, x_bf16, tokens: int, kv, **epilogue):
        _, wo = self.weights()
        return ops.linear(self.attend(x_bf16, tokens, kv), wo, self.to_out.bias, **epilogue)


class QKVAttention(torch.nn.Module):
    def __init__(self, num_heads, disable_self_attention: bool = False):
        super().__init__()
        self.num_heads = num_heads


#: key of the per-loop store {id(layer): encoder K/V buffer} in the sampling loop's static conditioning dict
TEXT_KV_KEY = "_xdb_unet_text_kv"


class ChanLayerNormGain(torch.nn.Module):
    """Parameter holder of the reference's ``LayerNorm(feats, dim)`` (attention.py:284-310): gain ``g`` only."""

    def __init__(self, feats, dim=-1):
        super().__init__()
        self.dim = dim
        self.g = torch.nn.Parameter(torch.ones(feats, *((1,) * (-dim - 1))))

    def forward(self, x):
        var = torch.var(x, dim=self.dim, unbiased=False, keepdim=True)
        mean = torch.mean(x, dim=self.dim, keepdim=True)
        return (x - mean) * (var + 1e-5).rsqrt() * self.g


class AttentionPooling(torch.nn.Module):
    """Imagen attention pooling (reference: attention.py:231-283): a class token (sequence mean + positional embedding)
    attends over [class token, sequence].  It only ever sees the text embeddings, which do not change during a sampling
    loop: it runs ONCE per loop (Unet.precompute_context) in plain fp32 tensor ops, never per timestep."""

    def __init__(self, num_heads, embed_dim):
        super().__init__()
        self.positional_embedding = torch.nn.Parameter(torch.randn(1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = torch.nn.Linear(embed_dim, embed_dim)
        self.q_proj = torch.nn.Linear(embed_dim, embed_dim)
        self.v_proj = torch.nn.Linear(embed_dim, embed_dim)
        self.num_heads = num_heads
        self.dim_per_head = embed_dim // self.num_heads

    def forward(self, x):
        bs, length, width = x.shape
        H, d = self.num_heads, self.dim_per_head
        cls = x.mean(dim=1, keepdim=True) + self.positional_embedding.to(x.dtype)
        xx = torch.cat([cls, x], dim=1)
        q = self.q_proj(cls).view(bs, 1, H, d).transpose(1, 2)                  # [bs, H, 1, d]
        k = self.k_proj(xx).view(bs, length + 1, H, d).transpose(1, 2)
        v = self.v_proj(xx).view(bs, length + 1, H, d).transpose(1, 2)
        w = torch.softmax((q @ k.transpose(-1, -2)).float() / math.sqrt(d), dim=-1)   # (d^-1/4)^2 on the logits
        # the reference returns `a.reshape(bs, -1, 1)` of a [bs*H, d, 1] tensor: channel = head * d + c
        return (w @ v).reshape(bs, H * d)


class SpatialCrossAttention(ContextBlock, Packed):
    """UNet attention block over pixels (reference: attention.py:20-141): GN32 -> qkv (1x1) -> per-head-interleaved
    softmax attention with ch^-1/4 on q and k -> proj_out (1x1, zero-init) -> residual.  With ``context_dim`` > 0
    (Imagen / GLIDE text conditioning) the text embeddings go through the context adapter, an optional channel
    LayerNorm and ``_encoder_kv`` (1x1), and the resulting keys / values are concatenated IN FRONT of the
    self-attention keys / values (attention.py:107-130,166-180).  The encoder half is timestep-invariant and is
    computed once per sampling loop (``encode_context``)."""

    def __init__(self, in_channels, context_dim=None, context_projection_input_dim=None,
                 context_projection_output_dim=None, heads=8, dim_head=64, dropout=0.0, context_layer_norm=False,
                 context_adapter=None, **kwargs):
        super().__init__()
        for k in ("pre_layer_norm", "post_layer_norm", "disable_self_attention", "is_video"):
            if kwargs.get(k):
                raise NotImplementedError(k)
        if context_projection_input_dim is not None and context_projection_output_dim is not None:
            raise NotImplementedError("_context_proj (factorised context projection)")
        context_dim = None if context_dim in (None, -1) else context_dim
        self._context_dim = context_dim
        self._channels = in_channels
        if dim_head == -1:
            self._num_heads = heads
        else:
            assert in_channels % dim_head == 0
            self._num_heads = in_channels // dim_head
        # sub-modules are registered in the reference's order (attention.py:61-98): LoRA files list their tensors in module
        # traversal order (xdiffusion_b200/lora.py)
        self._norm = torch.nn.GroupNorm(num_groups=32, num_channels=in_channels)
        if context_dim is not None:
            self._context_layer_norm = ChanLayerNormGain(context_dim, dim=-2) if context_layer_norm \
                else torch.nn.Identity()
        self._qkv = torch.nn.Conv1d(in_channels, in_channels * 3, 1)
        self._attention = QKVAttention(self._num_heads)
        if context_dim is not None:
            from ..utils import instantiate_from_config
            if not (context_adapter and "target" in context_adapter):
                raise NotImplementedError("SpatialCrossAttention(context_dim=...) without a context_adapter")
            self._context_adapter = instantiate_from_config(dict(context_adapter))
            self._encoder_kv = torch.nn.Conv1d(context_dim, in_channels * 2, 1)
        self._proj_out = zero_module(torch.nn.Conv1d(in_channels, in_channels, 1))

    def encode_context(self, context: Dict, tokens: int, kv=None):
        """Encoder keys / values of this block for the loop's conditioning (timestep-invariant): returns (or refreshes in
        place) the bf16 buffer [B, L + tokens, heads, 3, ch] whose first L rows carry [unused | k_enc | v_enc] per head;
        the forward appends the block's own [q | k | v] rows behind them every step."""
        ctx = self._context_layer_norm(self._context_adapter(context).float())       # [B, Cctx, L]
        B, Cc, L = ctx.shape
        heads = self._num_heads
        ch = self._channels // heads
        w = self.packed("wkv", (self._encoder_kv.weight,), lambda: bf16_weight(self._encoder_kv.weight))
        yb = ctx.permute(0, 2, 1).reshape(B * L, Cc).to(torch.bfloat16).contiguous()
        ekv = ops.linear(yb, w, self._encoder_kv.bias).view(B, L, heads, 2, ch)      # channel = head * 2ch + {k, v} * ch + c
        if kv is None or kv.shape != (B, L + tokens, heads, 3, ch):
            kv = torch.zeros((B, L + tokens, heads, 3, ch), device=ekv.device, dtype=torch.bfloat16)
        kv[:, :L, :, 1:].copy_(ekv)
        return kv

    def forward(self, x, context: Optional[Dict] = None, out=None, kv=None):
        """x bf16 NHWC [nimg, H, W, C] (a channel slice of a wider buffer is fine) -> same shape.
        ``kv``: the buffer of ``encode_context`` when the block has encoder context."""
        nimg, H, W, C = x.shape
        T, heads = H * W, self._num_heads
        ch = C // heads
        if ch != 64:
            raise NotImplementedError("attention head dim != 64")
        wq, wp = self.packed("w", (self._qkv.weight, self._proj_out.weight),
                             lambda: (bf16_weight(self._qkv.weight), bf16_weight(self._proj_out.weight)))
        xs = x.as_strided((nimg, T, C), (x.stride(0), x.stride(2), 1))
        n = ops.groupnorm(xs, self._norm.weight, self._norm.bias, eps=self._norm.eps)          # dense [nimg,T,C]
        qkv = ops.linear(n.view(nimg * T, C), wq, self._qkv.bias).view(nimg, T, heads, 3, ch)
        q, k, v = (qkv[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
        if self._context_dim is not None:
            # the sampling loop keeps one buffer per block in its static conditioning (filled on the first, eager step and
            # refreshed by Unet.precompute_context); a bare forward re-encodes the text every call, like the reference
            store = context.get(TEXT_KV_KEY) if context is not None else None
            if kv is None and store is not None:
                kv = store.get(id(self))
            if kv is None:
                if context is None or "text_embeddings" not in context:
                    raise RuntimeError("SpatialCrossAttention(context_dim > 0) needs text_embeddings in the context")
                kv = self.encode_context(context, T)
                if store is not None:
                    store[id(self)] = kv
            self.tokens_seen = T
            L = kv.shape[1] - T
            assert kv.shape[0] == nimg, "factorised (video) encoder context is not supported"
            torch.ops.xdb200.copy_rows(qkv.view(nimg, T, 3 * C), kv.view(nimg, L + T, 3 * C)[:, L:])
            k, v = (kv[:, :, :, i].permute(0, 2, 1, 3) for i in (1, 2))             # [nimg, heads, L + T, ch]
        a = ops.attention(q, k, v, 1.0 / math.sqrt(ch))             # (ch^-1/4)^2 on the logits
        a2 = a.permute(0, 2, 1, 3).reshape(nimg * T, C)
        if out is None:
            out = torch.empty((nimg, H, W, C), device=x.device, dtype=torch.bfloat16)
        x2 = x.as_strided((nimg * T, C), (x.stride(2), 1))
        o2 = out.as_strided((nimg * T, C), (out.stride(2), 1))
        ops.linear(a2, wp, self._proj_out.bias, residual=x2, out=o2, qstats=True)
        return out


class QKVAttentionWithRelativePosition(torch.nn.Module):
    """Holds the relative-position tables (reference: attention.py:490-549)."""

    def __init__(self, num_heads: int, max_relative_position: int, dim_head: int, sequence_length: int):
        super().__init__()
        self.num_heads = num_heads
        n = 2 * max_relative_position - 1
        self._max_relative_position = max_relative_position
        self._k_embeddings_table = torch.nn.Parameter(torch.randn(num_heads, n, dim_head) * dim_head ** -0.5)
        self._v_embeddings_table = torch.nn.Parameter(torch.randn(num_heads, n, dim_head) * dim_head ** -0.5)


class TemporalSelfAttention(ContextBlock, Packed):
    """Attention over frames at each pixel with relative-position logits (reference:
    attention.py:383-487, 551-676): GN over (C/32 x F) per pixel, no 1/sqrt(d) scale, and the
    reference's raw (B,H,L,D) -> (B,H*D,L) reinterpretation of the result."""

    def __init__(self, in_channels, temporal_sequence_length, max_relative_position, context_dim=None, heads=8,
                 dim_head=64, dropout=0.0, **kwargs):
        super().__init__()
        if context_dim not in (None, -1):
            raise NotImplementedError("temporal attention with encoder context")
        self._channels = in_channels
        self._num_heads = heads if dim_head == -1 else in_channels // dim_head
        self._dim_head = in_channels // self._num_heads
        self._length = temporal_sequence_length
        self._norm = torch.nn.GroupNorm(num_groups=32, num_channels=in_channels)
        self._qkv = torch.nn.Conv1d(in_channels, in_channels * 3, 1)
        self._attention = QKVAttentionWithRelativePosition(self._num_heads, max_relative_position, self._dim_head,
                                                           temporal_sequence_length)
        self._proj_out = zero_module(torch.nn.Conv1d(in_channels, in_channels, 1))

    def forward(self, x, frames: int, out=None):
        """x bf16 [B*F, H, W, C] (frames folded into the image count) -> same layout."""
        nimg, H, W, C = x.shape
        B, F, HW = nimg // frames, frames, H * W
        heads, d = self._num_heads, self._dim_head
        if d != 64:
            raise NotImplementedError("attention head dim != 64")
        # The logits q.k are NOT scaled by 1/sqrt(d) in the reference (attention.py:647): with |logit| ~ 10 the
        # softmax amplifies bf16 operand rounding (2^-9) into percent-level output error.  The qkv projection of
        # this block therefore runs in split precision on the same tensor-core kernel: activations and weights
        # as bf16 hi + lo pairs, x.W ~ x_hi.W_hi + x_lo.W_hi + x_hi.W_lo (K = 3C, error ~2^-16), fp32 q/k/v.
        def split_weights():
            w = self._qkv.weight.detach().reshape(3 * C, C).float()
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            return torch.cat([hi, hi, lo], 1).contiguous(), bf16_weight(self._proj_out.weight)
        wq3, wp = self.packed("w", (self._qkv.weight, self._proj_out.weight), split_weights)
        relk = self.packed("relk", (self._attention._k_embeddings_table,),
                           lambda: self._attention._k_embeddings_table.detach().float().contiguous())
        if relk.shape[1] != 2 * F - 1:
            raise NotImplementedError("temporal_sequence_length != max_relative_position")
        # GroupNorm over (C/32 x F) for every pixel of every clip: samples (b, hw), rows = frames
        rows = x.as_strided((nimg * HW, C), (x.stride(2), 1))
        n = torch.empty((nimg * HW, 2 * C), device=x.device, dtype=torch.bfloat16)        # [hi | lo]
        if not torch.ops.xdb200.groupnorm_frames_split(rows, self._norm.weight, self._norm.bias, self._norm.eps, B, F, HW, n):
            stats = torch.empty(B * HW * 64 * _lib.lib().xd_groupnorm_slabs(B * HW, F, C), device=x.device,
                                dtype=torch.float32)
            torch.ops.xdb200.groupnorm(rows, self._norm.weight, self._norm.bias, None, 1, self._norm.eps, 0, HW, B * HW,
                                       1, stats, n)
        qkv = ops.linear(n, wq3, self._qkv.bias, out_dtype=torch.float32, a2=n[:, :C])    # rows (b, f, hw) x 3C
        a = torch.empty((nimg * HW, C), device=x.device, dtype=torch.bfloat16)
        # ONE launch for all clips: kernel batch = clip, kernel "head" index = (pixel, head) -- the (pixel, head) pair has a
        # uniform stride (pixel: 3C = heads * 3d, head: 3d), `hpg = heads` recovers the true head for the rel-pos table
        # and the scrambled store, `o_gs = C` is the pixel's offset in the output.
        C3 = 3 * C
        q, k, v = (qkv.as_strided((B, HW * heads, F, d), (F * HW * C3, 3 * d, HW * C3, 1), qkv.storage_offset() + i * d)
                   for i in range(3))
        o = a.as_strided((B, HW * heads, F, d), (F * HW * C, d, HW * C, 1), a.storage_offset())
        ops.attention(q, k, v, 1.0, out=o, relk=relk, scramble=True, o_cs=1, hpg=heads, o_gs=C)
        if out is None:
            out = torch.empty((nimg, H, W, C), device=x.device, dtype=torch.bfloat16)
        o2 = out.as_strided((nimg * HW, C), (out.stride(2), 1))
        ops.linear(a, wp, self._proj_out.bias, residual=rows, out=o2, qstats=True)
        return out

"""Conditioning projections of the score networks, backed by the xdb200 kernels.

Class names, constructor arguments and parameter names follow the reference
(xdiffusion/layers/embedding.py) because YAML configs instantiate them by path and checkpoints
address their tensors by name; the forwards run on the C-ABI kernels.
"""
import math
from typing import Dict, List

import torch

from .. import ops
from .mlp import Mlp
from .utils import ContextBlock, Packed, bf16_weight


class ContextEmbedSequential(torch.nn.Sequential):
    """Container only: the owning network walks its children (reference: embedding.py:36-49)."""


def _sinusoid(t, freq, mode, max_time, clip, order, dtype=torch.bfloat16):
    out = torch.empty((t.shape[0], 2 * freq.shape[0]), device=t.device, dtype=dtype)
    torch.ops.xdb200.timestep_embed(t.contiguous(), freq, mode, float(max_time), float(clip[0]), float(clip[1]),
                                    order, out)
    return out


class _UNetTimeProjection(torch.nn.Module, Packed):
    """sinusoid(num_features) -> Linear -> SiLU -> Linear.  ``_projection.{1,3}`` hold the weights."""
    _mode = 1

    def __init__(self, num_features: int, time_embedding_mult: int, max_time: float = 1000.0, clip_min=-20,
                 clip_max=20, **kwargs):
        super().__init__()
        d = num_features * time_embedding_mult
        self._num_features, self._max_time, self._clip = num_features, max_time, (clip_min, clip_max)
        self._projection = torch.nn.Sequential(torch.nn.Identity(), torch.nn.Linear(num_features, d),
                                               torch.nn.SiLU(), torch.nn.Linear(d, d))

    def _freq(self, device):
        # same torch ops as the reference so the fp32 table is bit-identical (embedding.py:70-73)
        half = self._num_features // 2
        return self.packed("freq", (self._projection[1].weight,), lambda: torch.exp(
            torch.arange(half) * -(math.log(10000) / (half - 1))).to(device))

    def forward(self, timestep: torch.Tensor, **kwargs):
        l1, l3 = self._projection[1], self._projection[3]
        w1, w3 = self.packed("w", (l1.weight, l3.weight), lambda: (bf16_weight(l1.weight), bf16_weight(l3.weight)))
        if timestep.dtype not in (torch.int64, torch.float32):
            timestep = timestep.float()
        s = _sinusoid(timestep, self._freq(timestep.device), self._mode, self._max_time, self._clip, 0)
        h = ops.linear(s, w1, l1.bias, act=ops.ACT_SILU)
        return ops.linear(h, w3, l3.bias, out_dtype=torch.float32)


class TimestepEmbeddingProjection(_UNetTimeProjection):
    """reference: embedding.py:79-105 (its NaN check is a host sync and is not reproduced)."""
    _mode = 1


class InvCosTimestepEmbeddingProjection(_UNetTimeProjection):
    """input is logsnr: atan(exp(-clip(l)/2))/(pi/2) first (reference: embedding.py:108-143)."""
    _mode = 2


class DiTTimestepEmbedding(torch.nn.Module, Packed):
    """[cos|sin](256) -> Linear -> SiLU -> Linear (reference: embedding.py:325-343)."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int):
        super().__init__()
        self.mlp = torch.nn.Sequential(torch.nn.Linear(frequency_embedding_size, hidden_size, bias=True),
                                       torch.nn.SiLU(), torch.nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size

    def custom_initializer(self):
        torch.nn.init.normal_(self.mlp[0].weight, std=0.02)
        torch.nn.init.normal_(self.mlp[2].weight, std=0.02)

    def forward(self, timestep: torch.Tensor, **kwargs):
        l0, l2 = self.mlp[0], self.mlp[2]
        half = self.frequency_embedding_size // 2
        freq = self.packed("freq", (l0.weight,), lambda: torch.exp(
            -math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(l0.weight.device))
        w0, w2 = self.packed("w", (l0.weight, l2.weight), lambda: (bf16_weight(l0.weight), bf16_weight(l2.weight)))
        tab = (kwargs.get("context") or {}).get(TIMESTEP_TABLE_KEY)
        if tab is not None and tab[0] == id(self):       # the sampling loop evaluated this MLP for all of its timesteps
            return TimestepLookup(tab[1], tab[2], timestep.shape[0])
        if timestep.dtype not in (torch.int64, torch.float32):
            timestep = timestep.float()
        s = _sinusoid(timestep, freq, 0, 1.0, (0, 0), 1)
        h = ops.linear(s, w0, l0.bias, act=ops.ACT_SILU)
        return ops.linear(h, w2, l2.bias, out_dtype=torch.float32)


SILU_PREFIX = "_xdb_silu_"
ROWS_PREFIX = "_xdb_rows_"
TIMESTEP_TABLE_KEY = "_xdb_temb_table"      # context entry set by the sampling loop: (id(projection), table [N, D], loop index)


class TimestepLookup:
    """Deferred ``table[loop index]`` broadcast to ``rows`` rows: resolved by DiTCombineEmbeddngs in its one kernel."""

    def __init__(self, table, idx, rows):
        self.table, self.idx, self.rows = table, idx, rows


class LabelLookup:
    """Deferred ``embedding_table[labels]``: resolved by DiTCombineEmbeddngs in one fused kernel."""

    def __init__(self, table, labels, zero=False):
        self.table, self.labels, self.zero = table, labels, zero


class DiTLabelEmbedding(torch.nn.Module):
    """reference: embedding.py:346-382 (index ``num_classes`` is the null class)."""

    def __init__(self, num_classes, hidden_size, drop_prob: float = 0.0, unconditional_override: bool = False):
        super().__init__()
        self.embedding_table = torch.nn.Embedding(num_classes + 1, hidden_size)
        self.num_classes, self._unconditional_override, self._drop_prob = num_classes, unconditional_override, drop_prob
        torch.nn.init.normal_(self.embedding_table.weight, std=0.02)

    def forward(self, labels, **kwargs):
        if self._unconditional_override:
            labels = torch.zeros_like(labels) + self.num_classes
        if self._drop_prob not in (0.0, 1.0):
            raise NotImplementedError("random label dropping at sampling time")
        return LabelLookup(self.embedding_table.weight, labels, zero=self._drop_prob == 1.0)


class DiTCombineEmbeddngs(torch.nn.Module):
    """context[out] = sum of the source entries (reference: embedding.py:385-406)."""

    def __init__(self, output_context_key: str, source_context_keys: List[str], projections=None, **kwargs):
        super().__init__()
        self._output_context_key, self._source_context_keys = output_context_key, source_context_keys
        self._projections = projections

    def forward(self, context: Dict, **kwargs):
        vals = [context[k] for k in self._source_context_keys]
        lookups = [v for v in vals if isinstance(v, LabelLookup) and not v.zero]
        steps = [v for v in vals if isinstance(v, TimestepLookup)]
        if len(steps) == 1 and len(lookups) <= 1 and not any(torch.is_tensor(v) for v in vals):
            t, lk = steps[0], (lookups[0] if lookups else None)
            c = torch.empty((t.rows, t.table.shape[1]), device=t.table.device, dtype=torch.float32)
            silu = torch.empty(c.shape, device=c.device, dtype=torch.bfloat16)
            rows = torch.empty(t.rows, device=c.device, dtype=torch.int32) if lk else None
            torch.ops.xdb200.class_combine_step(lk.table if lk else None, lk.labels.contiguous() if lk else None, t.table, t.idx,
                                                t.rows, c, silu, rows)
            context[self._output_context_key] = c
            context[SILU_PREFIX + self._output_context_key] = silu          # bf16 SiLU(c): the input of the adaLN GEMMs
            if rows is not None:                                            # row of every image in a (label, step) table
                context[ROWS_PREFIX + self._output_context_key] = (rows, lk.table)
            return context
        dense = [v for v in vals if torch.is_tensor(v)]
        if len(dense) != 1 or len(lookups) > 1:
            raise NotImplementedError("DiTCombineEmbeddngs: expected one dense embedding (+ one label lookup)")
        out = dense[0]
        if lookups:
            c = torch.empty_like(out)
            torch.ops.xdb200.class_combine(lookups[0].table, lookups[0].labels.contiguous(), out, c, None)
            out = c
        context[self._output_context_key] = out
        return context


class RunProjection(torch.nn.Module):
    """context[out] = projections[key](context[in]) (reference: embedding.py:240-266)."""

    def __init__(self, input_context_key: str, output_context_key: str, projection_key: str, projections, **kwargs):
        super().__init__()
        self._input_context_key, self._output_context_key = input_context_key, output_context_key
        self._projection_key, self._projections = projection_key, projections

    def forward(self, context: Dict, device=None, **kwargs):
        assert self._input_context_key in context, \
            f"{self._input_context_key} not found for projection {self._projection_key}."
        context[self._output_context_key] = self._projections[self._projection_key](
            context[self._input_context_key], context=context, device=device)
        return context


class PooledTextEmbeddingsToTimestep(torch.nn.Module):
    """timestep_embedding += LayerNorm(Linear(AttentionPooling(LayerNorm(text_embeddings))))
    (reference: embedding.py:146-169).  The addend depends on the text only: ``pool`` computes it once per sampling
    loop (Unet.precompute_context stores it under POOLED_KEY); per timestep only the add runs, on the device kernel."""
    POOLED_KEY = "_xdb_pooled_text"

    def __init__(self, text_embedding_dim: int, time_embedding_dim: int, attention_pooling_heads: int, **kwargs):
        super().__init__()
        from .attention import AttentionPooling
        self._encoder_pooling = torch.nn.Sequential(
            torch.nn.LayerNorm(text_embedding_dim), AttentionPooling(attention_pooling_heads, text_embedding_dim),
            torch.nn.Linear(text_embedding_dim, time_embedding_dim), torch.nn.LayerNorm(time_embedding_dim))

    @torch.no_grad()
    def pool(self, context: Dict):
        return self._encoder_pooling(context["text_embeddings"].float()).contiguous()

    def forward(self, context: Dict, **kwargs):
        assert "text_embeddings" in context and "timestep_embedding" in context
        pooled = context.get(self.POOLED_KEY)
        if pooled is None:
            pooled = self.pool(context)
        t = context["timestep_embedding"].contiguous()
        out = torch.empty_like(t)
        torch.ops.xdb200.add_rows_periodic(t, pooled, t.shape[0], out)
        context["timestep_embedding"] = out
        return context


class ContextProjection(torch.nn.Module):
    """context[out] = Mlp(GELU-tanh)(context[in]) over (B, L, C) (reference: embedding.py:202-237)."""

    def __init__(self, input_context_key: str, output_context_key: str, in_features: int, hidden_features: int,
                 out_features: int, custom_initialization: bool = False, **kwargs):
        super().__init__()
        self._input_context_key, self._output_context_key = input_context_key, output_context_key
        self._custom_initialization = custom_initialization
        self.y_proj = Mlp(in_features=in_features, hidden_features=hidden_features, out_features=out_features)

    def custom_initializer(self):
        if self._custom_initialization:
            torch.nn.init.normal_(self.y_proj.fc1.weight, std=0.02)
            torch.nn.init.normal_(self.y_proj.fc2.weight, std=0.02)

    def forward(self, context: Dict, **kwargs):
        y = context[self._input_context_key]
        B, L, C = y.shape
        yb = torch.empty((B * L, C), device=y.device, dtype=torch.bfloat16)
        torch.ops.xdb200.act_cast(y.contiguous().view(B * L, C), ops.ACT_NONE, yb)
        context[self._output_context_key] = self.y_proj(yb).view(B, L, -1)
        return context


class PatchEmbed(torch.nn.Module, Packed):
    """Conv2d(k = s = patch) as im2col + GEMM (reference: embedding.py:409-508)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True, **kwargs):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.img_size = (img_size, img_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = torch.nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

    def forward(self, x, pos_embed=None):
        """x fp32 NCHW -> fp32 tokens [B*T, D] (+ pos_embed[T, D])."""
        B, C, H, W = x.shape
        assert (H, W) == self.img_size, f"Input size ({H},{W}) doesn't match model {self.img_size}."
        p = self.patch_size[0]
        w = self.packed("w", (self.proj.weight,), lambda: bf16_weight(self.proj.weight))
        cols = torch.empty((B * self.num_patches, C * p * p), device=x.device, dtype=torch.bfloat16)
        torch.ops.xdb200.patchify(x.contiguous(), p, cols)
        tok = ops.linear(cols, w, self.proj.bias, out_dtype=torch.float32)
        if pos_embed is not None:
            torch.ops.xdb200.add_rows_periodic(tok, pos_embed, self.num_patches, tok)
        return tok

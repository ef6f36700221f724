"""Shared helpers for the kernel-backed layers: a version-checked cache for repacked (bf16 /
kernel-layout) weights, zero-init, and the EinopsToAndFrom wrapper whose only job on this path is
to keep the reference's ``.fn.`` level in checkpoint keys (layers/utils.py:287-302)."""
import math

import numpy as np
import torch


class ContextBlock(torch.nn.Module):
    """Marker base: forward(x, context)."""


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class Packed:
    """Mixin: ``self.packed(key, params, build)`` returns ``build()`` cached until any of ``params``
    is written (``_version``) or moved (``data_ptr``) -- e.g. by load_state_dict / .to()."""

    def packed(self, key, params, build):
        cache = self.__dict__.setdefault("_pack_cache", {})
        sig = tuple((p.data_ptr(), p._version, p.device) for p in params)
        hit = cache.get(key)
        if hit is None or hit[0] != sig:
            with torch.no_grad():
                hit = (sig, build())
            cache[key] = hit
        return hit[1]


def bf16_weight(w):
    """[N, K...] fp32 parameter -> contiguous bf16 [N, K] (the GEMM's Wt operand)."""
    return w.detach().reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()


def pack_conv3x3(w, wskip=None):
    """[Cout, C, (1,) 3, 3] -> bf16 [Cout, 9*C (+Cs)] tap-major then channel; the optional 1x1 skip
    weights follow so that the skip projection accumulates in the same tensor-core tile."""
    co, c = w.shape[:2]
    p = w.detach().reshape(co, c, 3, 3).permute(0, 2, 3, 1).reshape(co, 9 * c)
    if wskip is not None:
        p = torch.cat([p, wskip.detach().reshape(co, -1)], 1)
    return p.to(torch.bfloat16).contiguous()


class EinopsToAndFrom(ContextBlock):
    def __init__(self, from_einops, to_einops, fn):
        super().__init__()
        self.from_einops, self.to_einops, self.fn = from_einops, to_einops, fn


def get_2d_sincos_pos_embed(embed_dim, grid_size, lewei_scale=1.0, base_size=16):
    """Fixed 2-D sin-cos table [grid*grid, D]: per token [sin(w om)|cos(w om)|sin(h om)|cos(h om)],
    om_k = 10000^(-k/(D/4)), positions arange(grid)/(grid/base_size)/lewei_scale
    (reference: layers/utils.py:188-258)."""
    if isinstance(lewei_scale, (tuple, list)):
        lewei_scale = lewei_scale[0]
    pos = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / lewei_scale
    quarter = embed_dim // 4
    omega = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)
    cols, rows = np.meshgrid(pos, pos)

    def enc(p):
        ang = np.einsum("m,d->md", p.reshape(-1), omega)
        return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)

    return np.concatenate([enc(cols), enc(rows)], axis=1)

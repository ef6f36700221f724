"""Imagen "efficient" UNet (the cascade's super-resolution stage) on the xdb200 kernels.

Drop-in for ``xdiffusion.score_networks.efficient_unet.Unet`` (reference: score_networks/efficient_unet.py:35-256): same
constructor (DotConfig), ``forward(x, context)`` and ``state_dict`` keys.  DBlocks (stride-2 conv as im2col + tcgen05 GEMM,
per-channel embedding bias, ResnetBlockEfficient chain, optional text cross-attention), UBlocks (the same, then nearest x2 +
conv3x3), skip concatenation, GroupNorm + SiLU + conv at the end.  The per-stage time-embedding linears run as ONE GEMM.
"""
from typing import Dict

import torch

from .. import ops
from ..layers.resnet import DBlock, UBlock
from ..layers.utils import Packed, bf16_weight
from ..utils import get_obj_from_str
from .dit import build_conditioning, run_custom_initializers


class Unet(torch.nn.Module, Packed):
    def __init__(self, config):
        super().__init__()
        self._config = config
        if config.is_learned_sigma:
            raise NotImplementedError("learned sigma")
        if config.is_class_conditional:
            raise NotImplementedError("class-conditional efficient UNet (label embedding)")
        self._is_learned_sigma = False
        self._output_channels = config.output_channels
        nf, mults = config.num_features, config.channel_multipliers
        time_emb_dim = nf * 4
        build_conditioning(self, config)
        # (the reference uses channels[0] = num_features * channel_multipliers[0] here and num_features for the first DBlock)
        self._initial_convolution = torch.nn.Conv2d(config.input_channels, nf * mults[0], 3, padding=1, bias=False)
        att_ds = [config.input_spatial_size // int(r) for r in config.attention.attention_resolutions]
        nres = config.num_resnet_blocks
        nres = nres if isinstance(nres, list) else [nres] * len(mults)
        layer = config.conditioning.context_transformer_layer.to_dict()
        att_type, att_kwargs = get_obj_from_str(layer["target"]), layer.get("params", {}) or {}

        def attention(ds):
            return (att_type, att_kwargs) if ds in att_ds else (None, {})

        chans, ch, ds = [nf], nf, 1
        self.downs = torch.nn.ModuleList([])
        for level, m in enumerate(mults):
            t, kw = attention(ds)
            self.downs.append(DBlock(dim_in=ch, dim_out=m * nf, dropout=config.dropout, num_resnet_blocks=nres[level],
                                     time_embedding_dim=time_emb_dim, downsample=True, attention_type=t,
                                     attention_kwargs=kw))
            ch = m * nf
            ds *= 2 if level != len(mults) - 1 else 1
            chans.append(ch)
        chans.pop()
        self.ups = torch.nn.ModuleList([])
        for level, m in list(enumerate(mults))[::-1]:
            t, kw = attention(ds)
            skip = chans.pop() if level != len(mults) - 1 else 0
            self.ups.append(UBlock(dim_in=ch + skip, dim_out=m * nf, dropout=config.dropout,
                                   num_resnet_blocks=nres[level] + 1, time_embedding_dim=time_emb_dim, upsample=True,
                                   attention_type=t, attention_kwargs=kw))
            ch = nf * m
            ds //= 2
        self.final_projection = torch.nn.Sequential(torch.nn.GroupNorm(32, nf), torch.nn.SiLU(),
                                                    torch.nn.Conv2d(nf, self._output_channels, 3, padding=1, bias=False))
        run_custom_initializers(self)

    def precompute_context(self, context):
        """Timestep-invariant text conditioning of the attention blocks (see score_networks/unet.py)."""
        from ..layers.attention import TEXT_KV_KEY, SpatialCrossAttention
        if "text_embeddings" not in context:
            return
        layers = {id(m): m for m in self.modules() if isinstance(m, SpatialCrossAttention) and m._context_dim is not None}
        if layers:
            store = context.setdefault(TEXT_KV_KEY, {})
            for key, kv in store.items():
                again = layers[key].encode_context(context, layers[key].tokens_seen, kv)
                assert again is kv

    def _emb_all(self):
        stages = list(self.downs) + list(self.ups)
        lins = [s.emb_linear() for s in stages]

        def build():
            offs, o = {}, 0
            for st, (w, _) in zip(stages, lins):
                offs[id(st)] = (o, w.shape[0])
                o += w.shape[0]
            return (torch.cat([bf16_weight(w) for w, _ in lins], 0), torch.cat([b.detach().float() for _, b in lins], 0), offs)
        return self.packed("emb_all", tuple(w for w, _ in lins) + tuple(b for _, b in lins), build)

    def forward(self, x, context: Dict):
        with ops.quad_stats():
            return self._forward(x, context)

    def _forward(self, x, context: Dict):
        context = context.copy()
        for ct in self._context_transformers:
            context = ct(context, device=x.device)
        temb = context["timestep_embedding"]
        B = x.shape[0]
        w, b, offs = self._emb_all()
        st = torch.empty(temb.shape, device=x.device, dtype=torch.bfloat16)
        torch.ops.xdb200.act_cast(temb.contiguous(), ops.ACT_SILU, st)
        emb = ops.linear(st, w, b, out_dtype=torch.float32)
        emb_of = lambda stage: emb[:, offs[id(stage)][0]: offs[id(stage)][0] + offs[id(stage)][1]]
        H, W = x.shape[2], x.shape[3]
        c0 = self._initial_convolution.weight.shape[0]
        h = torch.empty((B, H, W, c0), device=x.device, dtype=torch.bfloat16)
        torch.ops.xdb200.conv3x3_in(x.contiguous(), self._initial_convolution.weight, None, h)
        hs = []
        for stage in self.downs:
            h = stage(h, emb_of(stage), B, context)
            hs.append(h)
        hs.pop()
        for idx, stage in enumerate(self.ups):
            if idx > 0:                                   # torch.cat([h, skip], dim=1): two strided row copies
                skip = hs.pop()
                n, hh, ww, ch = h.shape
                cat = torch.empty((n, hh, ww, ch + skip.shape[3]), device=x.device, dtype=torch.bfloat16)
                torch.ops.xdb200.copy_rows(h.view(1, n * hh * ww, ch), cat.view(1, n * hh * ww, -1)[:, :, :ch])
                torch.ops.xdb200.copy_rows(skip.view(1, n * hh * ww, -1), cat.view(1, n * hh * ww, -1)[:, :, ch:])
                h = cat
            h = stage(h, emb_of(stage), B, context)
        gn = self.final_projection[0]
        hn = ops.groupnorm(h.view(B, -1, h.shape[3]), gn.weight, gn.bias, eps=gn.eps, silu=True).view(h.shape)
        y = torch.empty((B, self._output_channels, H, W), device=x.device, dtype=torch.float32)
        torch.ops.xdb200.conv3x3_out(hn, self.final_projection[2].weight.reshape(self._output_channels, -1, 3, 3), None, y)
        return y

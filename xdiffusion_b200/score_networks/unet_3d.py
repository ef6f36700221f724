"""Video Diffusion Models 3-D UNet (factorised space / time attention) on the xdb200 kernels.

Drop-in for ``xdiffusion.score_networks.unet_3d.Unet`` (reference: score_networks/unet_3d.py:27-353):
same constructor, ``forward(x [B,C,F,H,W], context)`` and ``state_dict`` keys.  Frames are folded
into the image count, so every (1,3,3) convolution, GroupNorm, per-frame spatial attention, pooling
and upsampling reuses the 2-D kernels; temporal attention runs over the frames of each pixel with
the relative-position logits and the reference's output reinterpretation.
"""
from typing import Dict

import torch

from .. import ops
from ..layers.attention import SpatialCrossAttention, TemporalSelfAttention
from ..layers.resnet_3d import ResnetBlockBigGAN3D
from ..layers.utils import EinopsToAndFrom
from ..utils import instantiate_partial_from_config
from .unet import Unet as Unet2D


class Unet(Unet2D):
    ResBlock = ResnetBlockBigGAN3D
    dims = 3

    @staticmethod
    def _conv(ci, co):
        return torch.nn.Conv3d(ci, co, kernel_size=(1, 3, 3), stride=(1, 1, 1), padding=(0, 1, 1), bias=False)

    @staticmethod
    def _res_kwargs(config):
        return {"mlp_layers": config.mlp_layers}

    @staticmethod
    def _attention_layers(config, ch, res):
        c = config.conditioning
        spatial = instantiate_partial_from_config(c.spatial_context_transformer_layer.to_dict())(
            in_channels=ch, context_projection_output_dim=res ** 2)
        temporal = instantiate_partial_from_config(c.temporal_context_transformer_layer.to_dict())(in_channels=ch)
        return [EinopsToAndFrom("b c f h w", "(b f) c h w", spatial),
                EinopsToAndFrom("b c f h w", "(b h w) c f", temporal)]

    def _to_nhwc_in(self, x):
        B, C, F, H, W = x.shape
        return x.permute(0, 2, 1, 3, 4).reshape(B * F, C, H, W), F

    def _from_nhwc_out(self, y, x):
        B, C, F, H, W = x.shape
        return y.view(B, F, -1, H, W).permute(0, 2, 1, 3, 4).contiguous()

    def _run_attention(self, layer, h, frames, out, context=None):
        fn = layer.fn
        if isinstance(fn, TemporalSelfAttention):
            return fn(h, frames, out=out)
        return fn(h, out=out)

    def _block_embeddings(self, temb, main):
        """[scale | shift] of every resblock.  The embedding Mlp stacks (2 small GEMMs per block, 44 launches of ~7 us per
        forward) depend on the timestep embedding only, so they all run up front on the conditioning (side) stream while
        ``main`` starts on the convolutions; a block waits for its own event just before it needs the rows (blocks are
        visited in registration order, so the last wait also joins the side stream)."""
        dev = temb.device
        side = torch.cuda.current_stream(dev)
        tb = torch.empty(temb.shape, device=dev, dtype=torch.bfloat16)
        torch.ops.xdb200.act_cast(temb.contiguous(), ops.ACT_NONE, tb)
        blocks = [m for m in self.modules() if isinstance(m, ResnetBlockBigGAN3D)]
        cache, events = {}, {}
        for blk in blocks:
            cache[id(blk)] = blk.embedding(tb)                  # kept alive until the forward returns
            ev = torch.cuda.Event()
            ev.record(side)
            events[id(blk)] = ev

        def emb_of(blk):
            ev = events.pop(id(blk), None)
            if ev is not None:
                main.wait_event(ev)
            return cache[id(blk)]
        return emb_of

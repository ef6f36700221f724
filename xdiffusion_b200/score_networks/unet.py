"""DDPM / IDDPM / GLIDE-style UNet score network on the xdb200 kernels.

Drop-in for ``xdiffusion.score_networks.unet.Unet`` (reference: score_networks/unet.py:35-299): same
constructor (DotConfig), ``forward(x, context)`` and ``state_dict`` keys.  Activations are NHWC bf16;
every skip connection lives in a pre-sized concat buffer that its producer (down path) and the
up-path producer write into directly, so ``torch.cat`` costs nothing; all 22 time-embedding linears
run as ONE GEMM per forward.
"""
from typing import Dict, List

import torch

from .. import ops
from ..layers.attention import TEXT_KV_KEY, SpatialCrossAttention
from ..layers.embedding import ContextEmbedSequential, PooledTextEmbeddingsToTimestep
from ..layers.resnet import Downsample, ResnetBlockBigGAN, Upsample
from ..layers.utils import Packed, bf16_weight
from ..utils import instantiate_partial_from_config
from .dit import build_conditioning, run_custom_initializers


class Unet(torch.nn.Module, Packed):
    ResBlock = ResnetBlockBigGAN
    dims = 2

    def __init__(self, config):
        super().__init__()
        self._config = config
        cin = config.input_channels
        self._output_channels = config.output_channels
        nf, mults = config.num_features, config.channel_multipliers
        if config.is_learned_sigma:
            raise NotImplementedError("learned sigma")
        if config.is_class_conditional:
            raise NotImplementedError("class-conditional UNet (label embedding)")
        if config.resnet_block_type != "biggan":
            raise NotImplementedError("resnet_block_type != 'biggan'")
        self._is_learned_sigma = False
        time_emb_dim = nf * 4
        build_conditioning(self, config)
        self._initial_convolution = self._conv(cin, nf)
        size = config.input_spatial_size
        size = size[1] if isinstance(size, list) else size
        att_res = config.attention.attention_resolutions if "attention" in config else config.attention_resolutions
        att_ds = [size // int(r) for r in att_res]
        nres = config.num_resnet_blocks
        nres = nres if isinstance(nres, list) else [nres] * len(mults)

        def res(ci, co):
            return self.ResBlock(dim_in=ci, time_emb_dim=time_emb_dim, dropout=config.dropout, dim_out=co,
                                 use_scale_shift_norm=config.use_scale_shift_norm, use_conv=config.resamp_with_conv,
                                 **self._res_kwargs(config))

        chans, ch, ds = [nf], nf, 1
        self.downs = torch.nn.ModuleList([])
        for level, m in enumerate(mults):
            for _ in range(nres[level]):
                layers = [res(ch, m * nf)]
                ch = m * nf
                if ds in att_ds:
                    layers += self._attention_layers(config, ch, size // ds)
                self.downs.append(ContextEmbedSequential(*layers))
                chans.append(ch)
            if level != len(mults) - 1:
                if config.resblock_updown:
                    raise NotImplementedError("resblock_updown=True")
                self.downs.append(ContextEmbedSequential(Downsample(ch, config.resamp_with_conv, dims=self.dims)))
                chans.append(ch)
                ds *= 2
        self.middle = ContextEmbedSequential(res(ch, ch), *self._attention_layers(config, ch, size // ds), res(ch, ch))
        self.ups = torch.nn.ModuleList([])
        for level, m in list(enumerate(mults))[::-1]:
            for i in range(nres[level] + 1):
                layers = [res(ch + chans.pop(), nf * m)]
                ch = nf * m
                if ds in att_ds:
                    layers += self._attention_layers(config, ch, size // ds)
                if level and i == nres[level]:
                    layers.append(Upsample(ch, config.resamp_with_conv, dims=self.dims))
                    ds //= 2
                self.ups.append(ContextEmbedSequential(*layers))
        self.final_projection = torch.nn.Sequential(torch.nn.GroupNorm(32, nf), torch.nn.SiLU(),
                                                    self._conv(nf, self._output_channels))
        run_custom_initializers(self)

    # ------------------------------------------------------------------ 2-D specifics (overridden for video)
    @staticmethod
    def _conv(ci, co):
        return torch.nn.Conv2d(ci, co, kernel_size=3, stride=1, padding=1, bias=False)

    @staticmethod
    def _res_kwargs(config):
        return {}

    @staticmethod
    def _attention_layers(config, ch, res):
        return [instantiate_partial_from_config(config.conditioning.context_transformer_layer.to_dict())(in_channels=ch)]

    def _to_nhwc_in(self, x):
        """fp32 NCHW input -> first conv -> bf16 NHWC [nimg,H,W,nf]; returns (nimg, frames)."""
        return x, 1

    def _from_nhwc_out(self, y, x):
        return y

    def _run_attention(self, layer, h, frames, out, context=None):
        return layer(h, context=context, out=out)

    # ------------------------------------------------------------------ timestep-invariant text conditioning
    def precompute_context(self, context):
        """Called by the sampling loop on its OWN static conditioning dict (at construction and after every in-place
        refresh).  Text-conditioned configs (Imagen base / GLIDE): the attention-pooled text embedding that is added to the
        timestep embedding and every attention block's encoder keys / values depend on the text only, so they are
        computed here, once per loop, into buffers that live in that dict (a captured CUDA graph keeps reading them; a
        refresh rewrites them in place)."""
        if "text_embeddings" not in context:
            return
        for ct in self._context_transformers:
            if isinstance(ct, PooledTextEmbeddingsToTimestep):
                new, old = ct.pool(context), context.get(ct.POOLED_KEY)
                if old is not None and old.shape == new.shape:
                    old.copy_(new)
                else:
                    context[ct.POOLED_KEY] = new
        layers = {id(m): m for m in self.modules() if isinstance(m, SpatialCrossAttention) and m._context_dim is not None}
        if layers:
            store = context.setdefault(TEXT_KV_KEY, {})
            for key, kv in store.items():               # blocks that already ran: re-encode into their buffers
                again = layers[key].encode_context(context, layers[key].tokens_seen, kv)
                assert again is kv

    # ------------------------------------------------------------------ forward
    def _emb_all(self):
        blocks = [m for m in self.modules() if isinstance(m, ResnetBlockBigGAN)]
        lins = [b.emb_linear() for b in blocks]
        params = tuple(w for w, _ in lins) + tuple(b for _, b in lins)

        def build():
            offs, o = {}, 0
            for blk, (w, _) in zip(blocks, lins):
                offs[id(blk)] = (o, w.shape[0])
                o += w.shape[0]
            return (torch.cat([bf16_weight(w) for w, _ in lins], 0),
                    torch.cat([b.detach().float() for _, b in lins], 0), offs)
        return self.packed("emb_all", params, build)

    def _side_stream(self, device):
        st = getattr(self, "_side", None)
        if st is None or st.device != device:
            st = torch.cuda.Stream(device=device)
            object.__setattr__(self, "_side", st)
        return st

    def _block_embeddings(self, temb, main):
        """[scale | shift] of every resblock from ONE GEMM over SiLU(temb).  Runs on the conditioning (side) stream; the
        accessor makes ``main`` wait for the rows the first time a block asks for them."""
        w, b, offs = self._emb_all()
        st = torch.empty(temb.shape, device=temb.device, dtype=torch.bfloat16)
        torch.ops.xdb200.act_cast(temb.contiguous(), ops.ACT_SILU, st)
        emb = ops.linear(st, w, b, out_dtype=torch.float32)
        ready = [torch.cuda.Event()]
        ready[0].record(torch.cuda.current_stream(temb.device))

        def emb_of(blk):
            if ready:
                main.wait_event(ready.pop())
            return emb[:, offs[id(blk)][0]: offs[id(blk)][0] + offs[id(blk)][1]]
        return emb_of

    def _run_entry(self, entry, h, emb_of, samples, frames, out, context=None):
        mods = list(entry)
        for j, layer in enumerate(mods):
            dst = out if j == len(mods) - 1 else None
            if isinstance(layer, ResnetBlockBigGAN):
                h = layer(h, (lambda blk=layer: emb_of(blk)), samples, out=dst)      # resolved right before the block's second GroupNorm
            elif isinstance(layer, (Downsample, Upsample)):
                h = layer(h, out=dst)
            else:
                h = self._run_attention(layer, h, frames, dst, context)
        return h

    def forward(self, x, context: Dict):
        with ops.quad_stats():                         # GroupNorm statistics ride along with the producing convs (ops.py)
            return self._forward(x, context)

    def _forward(self, x, context: Dict):
        context = context.copy()
        context["x"] = x
        # The conditioning branch (timestep / text embeddings, the resblocks' [scale | shift] rows) does not depend on the
        # activations: it runs on a side stream (a fork in the captured graph) next to the first convolution; the first block
        # that asks for its rows joins it, and everything the branch wrote into ``context`` is ordered before that point.
        main, side = torch.cuda.current_stream(x.device), self._side_stream(x.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for ct in self._context_transformers:
                context = ct(context, device=x.device)
            temb = context["timestep_embedding"]
            emb_of = self._block_embeddings(temb, main)
        samples = x.shape[0]
        x4, frames = self._to_nhwc_in(x)
        nimg, _, H, W = x4.shape
        dev = x.device

        # widths of the up path's `h` operand for every skip (pop order), to size the concat buffers
        up_res = [e[0] for e in self.ups]
        skip_w = []                                   # channels of skip j (down-path output j)
        first = self._initial_convolution.weight.shape[0]
        skip_w.append(first)
        for e in self.downs:
            r = [m for m in e if isinstance(m, ResnetBlockBigGAN)]
            skip_w.append(r[0].out_channels if r else skip_w[-1])
        n_skip = len(skip_w)
        h_w = [up_res[k].input_channels - skip_w[n_skip - 1 - k] for k in range(n_skip)]   # for up block k

        def cat_buffer(j, hh, ww):
            k = n_skip - 1 - j
            return torch.empty((nimg, hh, ww, h_w[k] + skip_w[j]), device=dev, dtype=torch.bfloat16)

        cats: List[torch.Tensor] = [None] * n_skip
        cats[0] = cat_buffer(0, H, W)
        h = cats[0][..., h_w[n_skip - 1]:]
        torch.ops.xdb200.conv3x3_in(x4.contiguous(), self._initial_convolution.weight, None, h)
        hh, ww = H, W
        taps = context.get("_taps")                    # debugging aid: {name: activation} (tools/taps_compare.py)
        for j, entry in enumerate(self.downs, start=1):
            if isinstance(entry[0], Downsample):
                hh, ww = hh // 2, ww // 2
            cats[j] = cat_buffer(j, hh, ww)
            h = self._run_entry(entry, h, emb_of, samples, frames, cats[j][..., h_w[n_skip - 1 - j]:], context)
            if taps is not None:
                taps[f"downs.{j - 1}"] = h.float().clone()
        h = self._run_entry(self.middle, h, emb_of, samples, frames, cats[n_skip - 1][..., :h_w[0]], context)
        if taps is not None:
            taps["middle"] = h.float().clone()
        for k, entry in enumerate(self.ups):
            if k + 1 < n_skip:
                nxt = cats[n_skip - 2 - k]
                dst = nxt[..., :h_w[k + 1]]
            else:
                dst = None
            h = self._run_entry(entry, cats[n_skip - 1 - k], emb_of, samples, frames, dst, context)
            if taps is not None:
                taps[f"ups.{k}"] = h.float().clone()
        gn = self.final_projection[0]
        hs = h.as_strided((samples, h.shape[0] // samples * h.shape[1] * h.shape[2], h.shape[3]),
                          (h.stride(0) * (h.shape[0] // samples), h.stride(2), 1))
        hn = ops.groupnorm(hs, gn.weight, gn.bias, eps=gn.eps, silu=True).view(h.shape)
        y = torch.empty((nimg, self._output_channels, H, W), device=dev, dtype=torch.float32)
        torch.ops.xdb200.conv3x3_out(hn, self.final_projection[2].weight.reshape(self._output_channels, -1, 3, 3),
                                     None, y)
        return self._from_nhwc_out(y, x)

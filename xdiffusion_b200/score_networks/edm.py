"""EDM preconditioning wrapper and the DDPM++ raw network (reference: score_networks/edm.py:16-238, 635-697).

``EDMPrecond`` keeps the reference constructor and ``forward(x, sigma, class_labels)``; the sampler does not call
``forward`` but ``raw`` + ``precond_scalars`` so that the output arithmetic runs inside the fused step kernel.  The four
scalars are evaluated with the reference's own fp32 tensor expressions (one value per step; on the host).

``SongUNet`` is built in its DDPM++ configuration (positional noise embedding, standard encoder / decoder, box resampling
filter, one 256-wide attention head: ``configs/image/mnist/edm.yaml``) on the NHWC bf16 kernels: every skip connection is
written by its producer straight into the decoder block's concat buffer, the ``affine(emb)`` terms of all blocks come from one
GEMM whose bias already holds ``conv0.bias`` (the noise embedding is one row for the whole batch, so the term is a conv
bias).  NCSN++ (Fourier embedding, residual encoder, [1, 3, 3, 1] filter) and ``DhariwalUNet`` (ADM) raise
NotImplementedError.  Any module with the raw signature ``model(x, noise_labels, class_labels=None)`` can be wrapped.
"""
import math

import torch

from .. import ops
from ..layers.edm import Conv2d, GroupNorm, Linear, UNetBlock
from ..layers.resnet import _as_samples
from ..layers.utils import Packed, bf16_weight
from ..utils import instantiate_from_config


class EDMPrecond(torch.nn.Module):
    def __init__(self, img_resolution, img_channels, label_dim=0, use_fp16=False, sigma_min=0, sigma_max=float("inf"),
                 sigma_data=0.5, model_type="DhariwalUNet", **model_kwargs):
        super().__init__()
        if use_fp16:
            raise NotImplementedError("use_fp16")
        self.img_resolution, self.img_channels, self.label_dim = img_resolution, img_channels, label_dim
        self.sigma_min, self.sigma_max, self.sigma_data = sigma_min, sigma_max, sigma_data
        self.model = instantiate_from_config(model_kwargs["model"])

    def precond_scalars(self, sigma):
        """(c_skip, c_out, c_in, c_noise) as Python floats holding the reference's fp32 values (edm.py:682-685)."""
        s = torch.as_tensor(sigma).to(torch.float32).reshape(-1, 1, 1, 1).cpu()
        sd = self.sigma_data
        c_skip = sd ** 2 / (s ** 2 + sd ** 2)
        c_out = s * sd / (s ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + s ** 2).sqrt()
        c_noise = s.log() / 4
        return float(c_skip), float(c_out), float(c_in), float(c_noise)

    def _labels(self, x, class_labels):
        if self.label_dim == 0:
            return None
        if class_labels is None:
            return torch.zeros([1, self.label_dim], device=x.device)
        return class_labels.to(torch.float32).reshape(-1, self.label_dim)

    def raw(self, xin, c_noise: float, class_labels=None):
        """F = model(c_in * x, c_noise): fp32 in, fp32 out."""
        noise_labels = torch.full((1,), c_noise, device=xin.device, dtype=torch.float32)
        out = self.model(xin, noise_labels, class_labels=self._labels(xin, class_labels))
        assert out.dtype == torch.float32
        return out.contiguous()

    def forward(self, x, sigma, class_labels=None, force_fp32=False, **model_kwargs):
        """D(x; sigma) = c_skip x + c_out F(c_in x; c_noise)  (one sigma for the batch, as the samplers call it)."""
        c_skip, c_out, c_in, c_noise = self.precond_scalars(sigma)
        x64 = x.to(torch.float64).contiguous()
        xin = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        torch.ops.xdb200.edm_prepare(x64, None, 0.0, None, c_in, xin)
        f = self.raw(xin, c_noise, class_labels)
        den = torch.empty_like(xin)
        torch.ops.xdb200.edm_step(2, x64, None, None, f, None, None, den, None, 1.0, 0.0, c_skip, c_out, 0.0)
        return den

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)


class SongUNet(torch.nn.Module, Packed):
    """DDPM++ (reference: score_networks/edm.py:16-238).  ``forward(x, noise_labels, class_labels)``: x fp32 NCHW,
    noise_labels fp32 [1] (one noise level for the batch, as the samplers call it) -> fp32 NCHW."""

    def __init__(self, img_resolution, in_channels, out_channels, label_dim=0, augment_dim=0, model_channels=128,
                 channel_mult=(1, 2, 2, 2), channel_mult_emb=4, num_blocks=4, attn_resolutions=(16,), dropout=0.10,
                 label_dropout=0, embedding_type="positional", channel_mult_noise=1, encoder_type="standard",
                 decoder_type="standard", resample_filter=(1, 1)):
        super().__init__()
        if embedding_type != "positional" or encoder_type != "standard" or decoder_type != "standard":
            raise NotImplementedError("NCSN++ variants (fourier embedding, skip / residual encoder, skip decoder)")
        if label_dim or augment_dim:
            raise NotImplementedError("class / augmentation labels")
        if in_channels != 1 or out_channels != 1:
            raise NotImplementedError("first / last convolution for other than one image channel")
        emb_channels, self.noise_channels = model_channels * channel_mult_emb, model_channels * channel_mult_noise
        kw = dict(emb_channels=emb_channels, num_heads=1, dropout=dropout, skip_scale=math.sqrt(0.5), eps=1e-6,
                  resample_filter=resample_filter, resample_proj=True, adaptive_scale=False)
        self.map_layer0 = Linear(self.noise_channels, emb_channels)
        self.map_layer1 = Linear(emb_channels, emb_channels)
        self.enc = torch.nn.ModuleDict()
        cout = in_channels
        for level, mult in enumerate(channel_mult):
            res = img_resolution >> level
            if level == 0:
                cin, cout = cout, model_channels
                self.enc[f"{res}x{res}_conv"] = Conv2d(cin, cout, 3)
            else:
                self.enc[f"{res}x{res}_down"] = UNetBlock(cout, cout, down=True, **kw)
            for idx in range(num_blocks):
                cin, cout = cout, model_channels * mult
                self.enc[f"{res}x{res}_block{idx}"] = UNetBlock(cin, cout, attention=res in attn_resolutions, **kw)
        skips = [b.out_channels for b in self.enc.values()]
        self.dec = torch.nn.ModuleDict()
        for level, mult in reversed(list(enumerate(channel_mult))):
            res = img_resolution >> level
            if level == len(channel_mult) - 1:
                self.dec[f"{res}x{res}_in0"] = UNetBlock(cout, cout, attention=True, **kw)
                self.dec[f"{res}x{res}_in1"] = UNetBlock(cout, cout, **kw)
            else:
                self.dec[f"{res}x{res}_up"] = UNetBlock(cout, cout, up=True, **kw)
            for idx in range(num_blocks + 1):
                cin, cout = cout + skips.pop(), model_channels * mult
                self.dec[f"{res}x{res}_block{idx}"] = UNetBlock(cin, cout, attention=idx == num_blocks and res in attn_resolutions, **kw)
            if level == 0:
                self.dec[f"{res}x{res}_aux_norm"] = GroupNorm(cout, eps=1e-6)
                self.dec[f"{res}x{res}_aux_conv"] = Conv2d(cout, out_channels, 3, init_weight=1e-5)

    def _blocks(self):
        return [b for b in list(self.enc.values()) + list(self.dec.values()) if isinstance(b, UNetBlock)]

    def _emb_bias(self, noise_labels):
        """conv0.bias + affine(emb) of every block from one GEMM: fp32 [1, sum Cout] and the column offsets."""
        blocks = self._blocks()
        params = tuple(p for b in blocks for p in (b.affine.weight, b.affine.bias, b.conv0.bias))
        params += (self.map_layer0.weight, self.map_layer1.weight)

        def build():
            offs, o = {}, 0
            for b in blocks:
                offs[id(b)] = (o, b.out_channels)
                o += b.out_channels
            half = self.noise_channels // 2
            freqs = torch.arange(half, dtype=torch.float32) / (half - 1)               # PositionalEmbedding(endpoint=True)
            freqs = ((1 / 10000) ** freqs).to(self.map_layer0.weight.device)
            return (torch.cat([bf16_weight(b.affine.weight) for b in blocks], 0),
                    torch.cat([(b.affine.bias + b.conv0.bias).detach().float() for b in blocks], 0), offs, freqs,
                    bf16_weight(self.map_layer0.weight), bf16_weight(self.map_layer1.weight))
        w, bias, offs, freqs, w0, w1 = self.packed("emb", params, build)
        e = torch.empty((1, self.noise_channels), device=noise_labels.device, dtype=torch.bfloat16)
        torch.ops.xdb200.timestep_embed(noise_labels.contiguous(), freqs, 0, 0.0, 0.0, 0.0, 0, e)   # [sin | cos] (:185-188)
        e = ops.linear(e, w0, self.map_layer0.bias, act=ops.ACT_SILU)
        e = ops.linear(e, w1, self.map_layer1.bias, act=ops.ACT_SILU)
        row = ops.linear(e, w, bias, out_dtype=torch.float32)[0]
        return lambda b: row[offs[id(b)][0]: offs[id(b)][0] + offs[id(b)][1]]

    def forward(self, x, noise_labels, class_labels=None, augment_labels=None):
        with ops.quad_stats():
            return self._forward(x, noise_labels)

    def _forward(self, x, noise_labels):
        if noise_labels.numel() != 1:
            raise NotImplementedError("one noise level per batch (the sampling path); got %d" % noise_labels.numel())
        B, _, H, W = x.shape
        dev = x.device
        bias_of = self._emb_bias(noise_labels.reshape(1).float())
        enc, dec = list(self.enc.values()), list(self.dec.values())
        # consumer of every skip: decoder blocks whose input is wider than the running activation take the newest skip
        res_of, r = [], H
        for m in enc:
            r = r // 2 if isinstance(m, UNetBlock) and m.down else r
            res_of.append(r)
        stack, ch, cats, consumer = list(range(len(enc))), enc[-1].out_channels, [None] * len(enc), {}
        for m in dec:
            if isinstance(m, UNetBlock):
                if m.in_channels != ch:
                    j = stack.pop()
                    cats[j] = torch.empty((B, res_of[j], res_of[j], m.in_channels), device=dev, dtype=torch.bfloat16)
                    consumer[id(m)] = (j, ch)
                ch = m.out_channels
        assert not stack
        width = {j: w for j, w in consumer.values()}

        def skip_slot(j):
            return cats[j][..., width[j]:]

        h = skip_slot(0)
        first = enc[0]
        torch.ops.xdb200.conv3x3_in(x.contiguous(), first.weight, first.bias, h)
        for j, m in enumerate(enc[1:], start=1):
            h = m(h, bias_of(m), B, out=skip_slot(j))
        blocks = [m for m in dec if isinstance(m, UNetBlock)]
        for i, m in enumerate(blocks):
            nxt = blocks[i + 1] if i + 1 < len(blocks) else None
            dst = None
            if nxt is not None and id(nxt) in consumer:
                j, w = consumer[id(nxt)]
                dst = cats[j][..., :w]
            src = cats[consumer[id(m)][0]] if id(m) in consumer else h
            h = m(src, bias_of(m), B, out=dst)
        norm, conv = dec[-2], dec[-1]
        hn = ops.groupnorm(_as_samples(h, B), norm.weight, norm.bias, eps=norm.eps, silu=True).view(h.shape)
        y = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
        torch.ops.xdb200.conv3x3_out(hn, conv.weight, conv.bias, y)
        return y

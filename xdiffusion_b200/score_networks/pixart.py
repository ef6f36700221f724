"""PixArt-alpha (adaLN-single + cross-attention) score network on the xdb200 kernels.

Drop-in for ``xdiffusion.score_networks.pixart.PixArtAlpha`` (reference: score_networks/pixart.py:24-359):
same constructor, ``forward(x, context)`` and ``state_dict`` keys.  The cross-attention K/V
projections and the ContextProjection of the text embeddings do not depend on the timestep; the
reference recomputes them every step (pixart.py:237-239,264-265); the sampling loop computes them once per
conditioning (``precompute_context``) and reuses them for every timestep (identical values, ~43% fewer FLOPs).
"""
import os
from typing import Dict

import torch

from .. import ops
from ..layers.attention import LastChannelCrossAttention, MultiHeadSelfAttention
from ..layers.embedding import TIMESTEP_TABLE_KEY, ContextProjection, PatchEmbed
from ..layers.mlp import Mlp
from ..layers.utils import Packed, bf16_weight, get_2d_sincos_pos_embed
from . import dit as _dit
from .dit import build_conditioning, run_custom_initializers

# Token rows from which a block runs on the fused DiT half-block kernels (csrc/dit_block.cu): LayerNorm + modulate + qkv +
# self-attention in one kernel, the cross-attention output projection + LayerNorm + modulate + fc1 + GELU + fc2 + gated
# residual in the other (the kernel's first stage is a 384 x 384 projection with a gated residual: here the cross-attention
# `to_out` with gate 1).  Per block 6 launches instead of 12; below the threshold one kernel per operator.
FUSED_MIN_ROWS = int(os.environ.get("XDB200_PIXART_FUSED_MIN_ROWS", "12288"))


class PixArtAlphaBlock(torch.nn.Module):
    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0, drop_path=0.0, window_size=0, use_rel_pos=False, **kw):
        super().__init__()
        if window_size or use_rel_pos:
            raise NotImplementedError("windowed / rel-pos PixArt attention")
        self.hidden_size = hidden_size
        self.attn = MultiHeadSelfAttention(hidden_size, num_heads=num_heads, qkv_bias=True)
        self.cross_attn = LastChannelCrossAttention(query_dim=hidden_size, context_dim=hidden_size, heads=num_heads,
                                                    dim_head=hidden_size // num_heads)
        self.mlp = Mlp(in_features=hidden_size, hidden_features=int(hidden_size * mlp_ratio))
        self.scale_shift_table = torch.nn.Parameter(torch.randn(6, hidden_size) / hidden_size ** 0.5)


class PixArtAlphaFinalLayer(torch.nn.Module):
    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.linear = torch.nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.scale_shift_table = torch.nn.Parameter(torch.randn(2, hidden_size) / hidden_size ** 0.5)
        self.out_channels = out_channels


class PixArtAlpha(torch.nn.Module, Packed):
    def __init__(self, config, **kwargs):
        super().__init__()
        self._config = config
        if config.is_learned_sigma:
            raise NotImplementedError("learned sigma")
        hidden, depth = config.hidden_size, config.depth
        self.in_channels = self.out_channels = config.input_channels
        self.patch_size, self.num_heads, self.hidden_size = config.patch_size, config.num_heads, hidden
        self.lewei_scale = (config.lewei_scale,)
        self.x_embedder = PatchEmbed(config.input_spatial_size, config.patch_size, config.input_channels, hidden)
        build_conditioning(self, config)
        self.base_size = config.input_spatial_size // config.patch_size
        self.register_buffer("pos_embed", torch.zeros(1, self.x_embedder.num_patches, hidden))
        self.t_block = torch.nn.Sequential(torch.nn.SiLU(), torch.nn.Linear(hidden, 6 * hidden, bias=True))
        wbi = config.window_block_indexes if "window_block_indexes" in config else []
        ws = config.window_size if "window_size" in config else 0
        self.blocks = torch.nn.ModuleList([
            PixArtAlphaBlock(hidden, config.num_heads, mlp_ratio=config.mlp_ratio,
                             window_size=ws if i in wbi else 0, use_rel_pos=config.use_rel_pos if i in wbi else False)
            for i in range(depth)])
        self.final_layer = PixArtAlphaFinalLayer(hidden, config.patch_size, self.out_channels)
        self.initialize_weights()

    def initialize_weights(self):
        """reference: pixart.py:311-359"""
        for m in self.modules():
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    torch.nn.init.constant_(m.bias, 0)
        grid = int(self.x_embedder.num_patches ** 0.5)
        self.pos_embed.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(
            self.hidden_size, grid, lewei_scale=self.lewei_scale, base_size=self.base_size)).float()[None])
        w = self.x_embedder.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        torch.nn.init.normal_(self._projections["timestep"].mlp[0].weight, std=0.02)
        torch.nn.init.normal_(self._projections["timestep"].mlp[2].weight, std=0.02)
        torch.nn.init.normal_(self.t_block[1].weight, std=0.02)
        for b in self.blocks:
            torch.nn.init.constant_(b.cross_attn.to_out.weight, 0)
            torch.nn.init.constant_(b.cross_attn.to_out.bias, 0)
        torch.nn.init.constant_(self.final_layer.linear.weight, 0)
        torch.nn.init.constant_(self.final_layer.linear.bias, 0)
        run_custom_initializers(self)

    def load_model_weights(self, state_dict: Dict):
        """pos_embed is a deterministic buffer and is dropped on load (reference: pixart.py:270-280)."""
        for key in ("pos_embed", "base_model.pos_embed", "model.pos_embed"):
            if key in state_dict:
                del state_dict[key]
                break
        self.load_state_dict(state_dict, strict=False)

    # ------------------------------------------------------------------ timestep-invariant conditioning
    KV_KEY = "_xdb_kv"

    def _compute_kv(self, context):
        """ContextProjection(text_embeddings) and every block's cross-attention K|V (bf16 [B, L, 2, H, d] each)."""
        c = dict(context)
        for ct in self._context_transformers:
            if isinstance(ct, ContextProjection):
                c = ct(c)
        y = c[self._config.context_key]                                   # bf16 [B, L, D]
        return [blk.cross_attn.project_context(y) for blk in self.blocks]

    def precompute_context(self, context):
        """Called by the sampling loop on its OWN static conditioning dict (at construction and after every in-place
        refresh): the K/V buffers live in that dict, so their lifetime is the loop's and a captured CUDA graph keeps
        reading valid memory; a refresh rewrites them in place.  Nothing is cached on the module -- a bare
        ``forward`` (no precomputed entry in the context) recomputes them, like the reference does every step."""
        if "context_key" not in self._config:
            return
        kvs = self._compute_kv(context)
        old = context.get(self.KV_KEY)
        if old is not None and len(old) == len(kvs) and old[0].shape == kvs[0].shape:
            for dst, new in zip(old, kvs):
                dst.copy_(new)
        else:
            context[self.KV_KEY] = kvs

    def _conditioning(self, context, B, device):
        """adaLN-single rows of every block, fp32 [depth, B, 7D] (six slices + a slice of exact ones), and of the final layer,
        fp32 [2, B, D], from the timestep embedding of the context."""
        D, depth = self.hidden_size, len(self.blocks)
        for ct in self._context_transformers:
            if not isinstance(ct, ContextProjection):
                context = ct(context=context, device=device)
        t = context["timestep_embedding"].contiguous()                    # fp32 [B, D]
        silu_t = torch.empty((B, D), device=device, dtype=torch.bfloat16)
        torch.ops.xdb200.act_cast(t, ops.ACT_SILU, silu_t)
        # adaLN-single: mod[n] = scale_shift_table[n] + t_block(t), six [B, D] slices per block.  A seventh slice of
        # exact ones rides along (zero weight rows, bias 1, zero table entries): the gate of the ungated cross-attention
        # residual for the fused kernel, at the same row pitch as the other slices and without a launch of its own.
        w_t, b_t = self.packed("t_block", (self.t_block[1].weight, self.t_block[1].bias), lambda: (
            torch.cat([bf16_weight(self.t_block[1].weight), torch.zeros((D, D), device=device, dtype=torch.bfloat16)], 0),
            torch.cat([self.t_block[1].bias.detach().float(), torch.ones(D, device=device)], 0)))
        t0 = ops.linear(silu_t, w_t, b_t, out_dtype=torch.float32)                        # [B, 7D]
        tables = self.packed("tables", tuple(b.scale_shift_table for b in self.blocks), lambda: torch.cat(
            [torch.stack([b.scale_shift_table.detach().reshape(-1) for b in self.blocks]).float(),
             torch.zeros((depth, D), device=device)], 1).contiguous())
        mod = torch.empty((depth, B, 7 * D), device=device, dtype=torch.float32)
        torch.ops.xdb200.add_table(t0, tables, mod)                   # table + t0 for all blocks
        fmod = torch.empty((2, B, D), device=device, dtype=torch.float32)
        torch.ops.xdb200.add_table(t, self.final_layer.scale_shift_table.detach().float().contiguous(), fmod)
        return mod, fmod

    @torch.no_grad()
    def timestep_table(self, timesteps, context=None):
        """(id(self), table): the adaLN-single rows of all blocks and the final layer for every timestep of a sampling loop,
        fp32 [N, depth * 7D + 2D] (row = loop index), when the conditioning depends on the timestep only -- checked against
        the loop's own context at one timestep; None otherwise (the rows are then computed per step)."""
        if context is None or timesteps.dim() != 1:
            return None
        N, dev = timesteps.shape[0], timesteps.device
        D, depth = self.hidden_size, len(self.blocks)
        try:
            probe = {k: v for k, v in context.items()}
            B = next(v.shape[0] for v in probe.values() if torch.is_tensor(v) and v.dim() > 0)
            probe["timestep"] = timesteps[-1:].expand(B).contiguous()
            mod_b, fmod_b = self._conditioning(probe, B, dev)
            generic = {"timestep": timesteps.contiguous()}
            if "classes" in context:
                generic["classes"] = torch.zeros(N, dtype=torch.long, device=dev)
            mod_n, fmod_n = self._conditioning(generic, N, dev)
        except Exception:  # noqa: BLE001  (a conditioning head that needs more than the timestep)
            return None
        if not (torch.equal(mod_b, mod_n[:, -1:].expand_as(mod_b)) and torch.equal(fmod_b, fmod_n[:, -1:].expand_as(fmod_b))):
            return None
        table = torch.cat([mod_n.permute(1, 0, 2).reshape(N, depth * 7 * D), fmod_n.permute(1, 0, 2).reshape(N, 2 * D)], 1)
        return id(self), table.contiguous()

    def forward(self, x, context: Dict, **kwargs):
        context = context.copy()
        kvs = context.get(self.KV_KEY)
        if kvs is None and "context_key" in self._config:
            kvs = self._compute_kv(context)
        B, D, T = x.shape[0], self.hidden_size, self.x_embedder.num_patches
        depth = len(self.blocks)
        # Two independent branches before the first block, as in the DiT: (a) conditioning -> adaLN-single rows on a side
        # stream (a fork / join in the captured graph), (b) patch embedding on the main stream.
        main, side = torch.cuda.current_stream(x.device), _dit.DiT._side_stream(self, x.device)
        side.wait_stream(main)
        tab = context.get(TIMESTEP_TABLE_KEY)
        with torch.cuda.stream(side):
            if tab is not None and tab[0] == id(self):
                # the loop evaluated the whole timestep-only conditioning for all of its timesteps (timestep_table): this
                # step's row is copied out and every image reads it (row stride 0)
                row = torch.empty(tab[1].shape[1], device=x.device, dtype=torch.float32)
                torch.ops.xdb200.gather_row(tab[1], tab[2], row)
                mod = row[:depth * 7 * D].view(depth, 1, 7 * D).expand(depth, B, 7 * D)
                fmod = row[depth * 7 * D:].view(2, 1, D).expand(2, B, D)
            else:
                mod, fmod = self._conditioning(context, B, x.device)
        h = self.x_embedder(x, self.pos_embed[0])                         # fp32 [B*T, D]
        main.wait_stream(side)
        fused = (_dit.FUSED_BLOCK and kvs is not None and D == 384 and T == 16 and self.num_heads * 64 == D
                 and ops.MATMUL_BACKEND == "tc" and self.blocks[0].mlp.act == ops.ACT_GELU
                 and self.blocks[0].mlp.fc1.out_features == 4 * D)
        if ops.BATCH_DEPENDENT_PATHS and B * T < FUSED_MIN_ROWS:
            fused = False                                 # (never when bit-exact batch independence is requested)
        if fused:
            stats = torch.empty((B * T, 2), device=x.device, dtype=torch.float32)       # (mean, rstd) of h, block n -> n + 1
        for n, blk in enumerate(self.blocks):
            s1, sc1, g1, s2, sc2, g2, one = (mod[n, :, i * D:(i + 1) * D] for i in range(7))
            if fused:
                wh, bh = blk.attn.head_packed()
                o = torch.empty((B * T, D), device=x.device, dtype=torch.bfloat16)
                torch.ops.xdb200.dit_attn(h, stats if n > 0 else None, s1, sc1, T, 1e-6, wh, bh, self.num_heads,
                                          blk.attn.scale, o)
                _, wp = blk.attn.weights()
                ops.linear(o, wp, blk.attn.proj.bias, gate=g1, gate_rows=T, residual=h, out=h)
                hb = torch.empty((B * T, D), device=x.device, dtype=torch.bfloat16)
                torch.ops.xdb200.act_cast(h, ops.ACT_NONE, hb)
                a = blk.cross_attn.attend(hb, T, kvs[n])
                _, wo = blk.cross_attn.weights()
                w1, w2 = blk.mlp.weights()
                torch.ops.xdb200.dit_proj_mlp(a, wo, blk.cross_attn.to_out.bias, w1, blk.mlp.fc1.bias, w2, blk.mlp.fc2.bias,
                                              h, h, one, s2, sc2, g2, T, 1e-6, stats, 1)
                continue
            a = ops.layernorm_modulate(h, s1, sc1, T)
            blk.attn(a, T, gate=g1, gate_rows=T, residual=h, out=h)
            if kvs is not None:
                hb = torch.empty((B * T, D), device=x.device, dtype=torch.bfloat16)
                torch.ops.xdb200.act_cast(h, ops.ACT_NONE, hb)
                blk.cross_attn(hb, T, kvs[n], residual=h, out=h)          # x += cross_attn(x, y): no norm, no gate
            a = ops.layernorm_modulate(h, s2, sc2, T)
            blk.mlp(a, gate=g2, gate_rows=T, residual=h, out=h)
        a = ops.layernorm_modulate(h, fmod[0], fmod[1], T)
        w_lin = self.packed("final", (self.final_layer.linear.weight,),
                            lambda: bf16_weight(self.final_layer.linear.weight))
        y = ops.linear(a, w_lin, self.final_layer.linear.bias, out_dtype=torch.float32)
        out = torch.empty((B, self.out_channels, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        torch.ops.xdb200.unpatchify(y, self.patch_size, out)
        return out

"""DiT (adaLN-Zero) score network on the xdb200 kernels.

Drop-in for ``xdiffusion.score_networks.dit.DiT`` (reference: score_networks/dit.py:77-229): same
constructor (a DotConfig), same ``forward(x, context) -> Tensor`` and identical ``state_dict`` keys.
Per forward: patchify+GEMM, ONE batched adaLN GEMM for all blocks (the conditioning vector is
block-independent), then per block LayerNorm+modulate -> QKV GEMM -> fused attention -> proj GEMM
with gate*+residual epilogue -> LayerNorm+modulate -> fc1 GEMM (+GELU) -> fc2 GEMM (gate*+residual);
from 768 images per GPU two fused tcgen05 kernels per block instead (csrc/dit_block.cu).
The token residual stream stays fp32; GEMM operands are bf16 with fp32 accumulation in TMEM.
Inside a sampling loop the conditioning does not run per step at all: ``timestep_table`` evaluates the
timestep MLP for all N timesteps and the adaLN modulation for every (label, timestep) pair once, and
the fused kernels read each image's row of that table through an index.
"""
import os
from typing import Dict

import torch

from .. import ops
from ..layers.attention import MultiHeadSelfAttention
from ..layers.embedding import (ROWS_PREFIX, SILU_PREFIX, TIMESTEP_TABLE_KEY, DiTLabelEmbedding, DiTTimestepEmbedding,
                                PatchEmbed)
from ..layers.mlp import Mlp
from ..layers.utils import Packed, bf16_weight, get_2d_sincos_pos_embed
from ..utils import instantiate_from_config, instantiate_partial_from_config


# Fused half-block kernels (csrc/dit_block.cu); XDB200_DIT_FUSED=0 runs one kernel per operator (round-1 path).
FUSED_BLOCK = os.environ.get("XDB200_DIT_FUSED", "1") == "1"
# CTA pairs per 256-row tile of the fused MLP kernel: 1 = never split (default: measured, the distributed-shared-memory
# reduction costs what the shorter MLP loop saves, profiles/README.md), 0 = chosen by the library, 2..4 = forced
MLP_SPLIT = int(os.environ.get("XDB200_DIT_MLP_SPLIT", "1"))
# The fused kernels walk one 256-row tile per CTA pair through a whole half-block: ~43 us however few tiles there are.
# Below this many token rows (768 images; e.g. the 128-image shards of a batch sharded 8 ways) one kernel per operator,
# each spread over all SMs along N, is faster (B = 128: 0.56 vs 0.71 ms per timestep; B = 1024: 1.42 vs 1.12).
FUSED_MIN_ROWS = int(os.environ.get("XDB200_DIT_FUSED_MIN_ROWS", "12288"))
# Budget for the per-loop (label, timestep) modulation table of the fused path (fp32 [(classes + 1) * N, depth * 6D + 2D]:
# 1.2 GB for 10 classes x 1000 steps); above it, or with XDB200_DIT_MOD_TABLE_MB=0, the adaLN GEMM runs every step.
MOD_TABLE_BYTES = int(os.environ.get("XDB200_DIT_MOD_TABLE_MB", "4096")) << 20


class DiTBlock(torch.nn.Module):
    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.attn = MultiHeadSelfAttention(hidden_size, num_heads=num_heads, qkv_bias=True)
        self.mlp = Mlp(in_features=hidden_size, hidden_features=int(hidden_size * mlp_ratio))
        self.adaLN_modulation = torch.nn.Sequential(torch.nn.SiLU(),
                                                    torch.nn.Linear(hidden_size, 6 * hidden_size, bias=True))


class FinalLayer(torch.nn.Module):
    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.linear = torch.nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = torch.nn.Sequential(torch.nn.SiLU(),
                                                    torch.nn.Linear(hidden_size, 2 * hidden_size, bias=True))


def build_conditioning(net, config):
    """Projections + context-transformer head, shared by DiT / PixArt / UNets
    (reference: score_networks/dit.py:105-130, unet.py:73-98)."""
    net._projections = torch.nn.ModuleDict()
    for name in config.conditioning.signals:
        net._projections[name] = instantiate_from_config(config.conditioning.projections[name].to_dict())
    head = config.conditioning.context_transformer_head
    head = head if isinstance(head, list) else [head.to_dict()]
    net._context_transformers = torch.nn.ModuleList(
        [instantiate_partial_from_config(c)(projections=net._projections) for c in head])


def run_custom_initializers(net):
    for m in net.modules():
        if hasattr(m, "custom_initializer"):
            m.custom_initializer()


class DiT(torch.nn.Module, Packed):
    def __init__(self, config):
        super().__init__()
        hidden, depth = config.hidden_size, config.depth
        self.learn_sigma = config.is_learned_sigma
        if self.learn_sigma:
            raise NotImplementedError("learned sigma")
        self.in_channels = config.input_channels
        self.out_channels = config.input_channels
        self.patch_size, self.num_heads, self.hidden_size = config.patch_size, config.num_heads, hidden
        self.x_embedder = PatchEmbed(config.input_spatial_size, config.patch_size, config.input_channels, hidden)
        build_conditioning(self, config)
        self.pos_embed = torch.nn.Parameter(torch.zeros(1, self.x_embedder.num_patches, hidden), requires_grad=False)
        self.blocks = torch.nn.ModuleList([DiTBlock(hidden, config.num_heads, config.mlp_ratio) for _ in range(depth)])
        self.final_layer = FinalLayer(hidden, config.patch_size, self.out_channels)
        self.initialize_weights()

    def initialize_weights(self):
        """Same scheme as the reference (dit.py:148-185), incl. the zero-initialised adaLN / final layers."""
        for m in self.modules():
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    torch.nn.init.constant_(m.bias, 0)
        grid = int(self.x_embedder.num_patches ** 0.5)
        self.pos_embed.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(self.hidden_size, grid)).float()[None])
        w = self.x_embedder.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        torch.nn.init.constant_(self.x_embedder.proj.bias, 0)
        for lin in [b.adaLN_modulation[-1] for b in self.blocks] + [self.final_layer.adaLN_modulation[-1],
                                                                   self.final_layer.linear]:
            torch.nn.init.constant_(lin.weight, 0)
            torch.nn.init.constant_(lin.bias, 0)
        run_custom_initializers(self)

    def timestep_table(self, timesteps, context=None):
        """(id(projection), table): the timestep MLP for every timestep of a sampling loop, [N, D] fp32 (row = loop index).
        The loop hands it back through ``context[TIMESTEP_TABLE_KEY]`` with its device-resident loop index, and the per-step
        conditioning chain (sinusoid, two GEMMs, combine, SiLU cast: five launches) becomes one kernel in front of the
        adaLN GEMM.  None when the conditioning head is not the plain DiT one."""
        mods = [m for m in self._projections.values() if isinstance(m, DiTTimestepEmbedding)]
        if len(mods) != 1 or timesteps.dtype not in (torch.int64, torch.float32):
            return None
        with torch.no_grad():
            temb = mods[0](timesteps.contiguous()).contiguous()
            return id(mods[0]), temb, self._modulation_table(temb)

    def _modulation_table(self, temb):
        """adaLN_modulation(SiLU(t_emb[step] + y_emb[label])) of every block (+ the final layer) for every (label, step) pair:
        fp32 [(classes + 1) * N, depth * 6D + 2D], row = label * N + step, or None (no class conditioning / over budget).
        The values are those the per-step path computes (same kernels, same operand order), so both paths agree bit for bit."""
        labels = [m for m in self._projections.values() if isinstance(m, DiTLabelEmbedding)]
        if len(labels) != 1 or labels[0]._drop_prob != 0.0:
            return None
        emb = labels[0].embedding_table.weight.detach().float()
        w_ada, b_ada = self._adaln()
        N, D = temb.shape
        if emb.shape[0] * N * w_ada.shape[0] * 4 > MOD_TABLE_BYTES or emb.device != temb.device:
            return None
        c_all = (emb[:, None, :] + temb[None, :, :]).reshape(-1, D).contiguous()          # table[label] + temb: the combine's order
        silu_all = torch.empty(c_all.shape, device=c_all.device, dtype=torch.bfloat16)
        torch.ops.xdb200.act_cast(c_all, ops.ACT_SILU, silu_all)
        return emb, ops.linear(silu_all, w_ada, b_ada, out_dtype=torch.float32)

    # ------------------------------------------------------------------ kernels
    def _side_stream(self, device):
        st = getattr(self, "_side", None)
        if st is None or st.device != device:
            st = torch.cuda.Stream(device=device)
            object.__setattr__(self, "_side", st)
        return st

    def _adaln(self):
        """All adaLN linears stacked: one [sum(6D..)+2D, D] GEMM per step instead of depth+1."""
        lins = [b.adaLN_modulation[1] for b in self.blocks] + [self.final_layer.adaLN_modulation[1]]
        params = tuple(l.weight for l in lins) + tuple(l.bias for l in lins)
        return self.packed("adaln", params, lambda: (
            torch.cat([bf16_weight(l.weight) for l in lins], 0), torch.cat([l.bias.detach().float() for l in lins], 0)))

    def forward(self, x, context: Dict):
        context = context.copy()
        B, D, T = x.shape[0], self.hidden_size, self.x_embedder.num_patches
        w_ada, b_ada = self._adaln()
        # Two independent branches before the first block: (a) conditioning -> adaLN GEMM (~60 us of small kernels),
        # (b) patch embedding.  (a) runs on a side stream (a fork/join in the captured graph); its outputs are
        # allocated on the main stream so that the caching allocator never recycles them early.
        main = torch.cuda.current_stream(x.device)
        side = self._side_stream(x.device)
        fused = FUSED_BLOCK and D == 384 and ops.MATMUL_BACKEND == "tc"
        if ops.BATCH_DEPENDENT_PATHS and B * T < FUSED_MIN_ROWS:
            fused = False                                 # (never when bit-exact batch independence is requested)
        fused_attn = fused and T == 16 and self.num_heads * 64 == D and self.blocks[0].attn.qkv.bias is not None
        silu_c = torch.empty((B, D), device=x.device, dtype=torch.bfloat16)
        base = len(self.blocks) * 6 * D
        tab = context.get(TIMESTEP_TABLE_KEY)
        modtab = tab[3] if tab is not None and len(tab) > 3 else None            # per-loop (label, step) modulation table
        mod = mod_rows = None
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for ct in self._context_transformers:
                context = ct(context, device=x.device)
            pre = context.get(SILU_PREFIX + "timestep_embedding")     # set when the loop's timestep table is in use
            if pre is not None:
                silu_c = pre
            else:
                c = context["timestep_embedding"]                      # fp32 [B, D]
                torch.ops.xdb200.act_cast(c.contiguous(), ops.ACT_SILU, silu_c)
            rows = context.get(ROWS_PREFIX + "timestep_embedding")
            if (modtab is not None and rows is not None and fused_attn and rows[1].data_ptr() == modtab[0].data_ptr()
                    and self.blocks[0].mlp.act == ops.ACT_GELU):
                # every image reads its (label, step) row of the table: no adaLN GEMM, at most classes + 1 distinct rows
                # per block and step (L2-resident); only the final layer's two slices are materialised per image
                mod_rows, mod = rows[0], modtab[1]
                fin = ops.linear(silu_c, w_ada[base:], b_ada[base:], out_dtype=torch.float32)      # [B, 2D]
            else:
                mod = torch.empty((B, w_ada.shape[0]), device=x.device, dtype=torch.float32)       # [B, depth*6D + 2D]
                ops.linear(silu_c, w_ada, b_ada, out=mod)
                fin = mod[:, base:]
        h = self.x_embedder(x, self.pos_embed[0])                      # fp32 [B*T, D]
        main.wait_stream(side)
        # (mean, rstd) of every row of h: emitted by the fused MLP kernel of block n, consumed by the LayerNorm of block n + 1
        stats = torch.empty((B * T, 2), device=x.device, dtype=torch.float32) if fused else None
        # The fused MLP kernel reads h and writes h_next.  In place by default: a second 25 MB fp32 stream (batch 1024) does
        # not fit the 126 MB L2 next to h, the attention output and the weights, and costs 9 % of the step (838 vs 914
        # img/s).  Only the hidden-split kernels (XDB200_DIT_MLP_SPLIT != 1) need distinct buffers.
        h_next = h if MLP_SPLIT == 1 or not fused else torch.empty_like(h)
        for n, blk in enumerate(self.blocks):
            m = mod[:, n * 6 * D:(n + 1) * 6 * D]
            s1, sc1, g1, s2, sc2, g2 = (m[:, i * D:(i + 1) * D] for i in range(6))
            if fused and blk.mlp.act == ops.ACT_GELU:
                # two kernels per block (csrc/dit_block.cu): LayerNorm-modulate + qkv + attention, then proj + gated
                # residual + LayerNorm-modulate + fc1 + GELU + fc2 + gated residual; neither the [B*T, 3D] qkv nor the
                # [B*T, 4D] MLP activation leaves the SM
                _, wp = blk.attn.weights()
                w1, w2 = blk.mlp.weights()
                if mod_rows is not None:
                    wh, bh = blk.attn.head_packed()
                    o = torch.empty((B * T, D), device=x.device, dtype=torch.bfloat16)
                    torch.ops.xdb200.dit_attn_rows(h, stats if n > 0 else None, s1, sc1, T, 1e-6, wh, bh, self.num_heads,
                                                   blk.attn.scale, mod_rows, o)
                    torch.ops.xdb200.dit_proj_mlp_rows(o, wp, blk.attn.proj.bias, w1, blk.mlp.fc1.bias, w2, blk.mlp.fc2.bias, h,
                                                       h_next, g1, s2, sc2, g2, T, 1e-6, stats, MLP_SPLIT, mod_rows)
                    h, h_next = h_next, h
                    continue
                if fused_attn:
                    wh, bh = blk.attn.head_packed()
                    o = torch.empty((B * T, D), device=x.device, dtype=torch.bfloat16)
                    torch.ops.xdb200.dit_attn(h, stats if n > 0 else None, s1, sc1, T, 1e-6, wh, bh, self.num_heads,
                                              blk.attn.scale, o)
                else:
                    o = blk.attn.attend(h, T, ln=(s1, sc1, T))
                torch.ops.xdb200.dit_proj_mlp(o, wp, blk.attn.proj.bias, w1, blk.mlp.fc1.bias, w2, blk.mlp.fc2.bias, h, h_next,
                                              g1, s2, sc2, g2, T, 1e-6, stats, MLP_SPLIT)
                h, h_next = h_next, h
                continue
            # ln=: LayerNorm + modulate feeding qkv / fc1 (layernorm_modulate kernel; fused into the GEMM with XDB200_LN_FUSED=1)
            blk.attn(h, T, ln=(s1, sc1, T), gate=g1, gate_rows=T, residual=h, out=h)    # h += g1 * attn(modulate(norm(h)))
            blk.mlp(h, ln=(s2, sc2, T), gate=g2, gate_rows=T, residual=h, out=h)        # h += g2 * mlp(modulate(norm(h)))
        a = ops.layernorm_modulate(h, fin[:, :D], fin[:, D:2 * D], T)
        w_lin = self.packed("final", (self.final_layer.linear.weight,),
                            lambda: bf16_weight(self.final_layer.linear.weight))
        y = ops.linear(a, w_lin, self.final_layer.linear.bias, out_dtype=torch.float32)
        out = torch.empty((B, self.out_channels, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        torch.ops.xdb200.unpatchify(y, self.patch_size, out)
        return out

"""Score-network forwards (oracle).  TEST INFRASTRUCTURE ONLY.

Functional fp32 restatements driven by a state dict (keys relative to
``_score_network.``) and the ``score_network.params`` dict of the YAML config.

  unet_forward     score_networks/unet.py:265-299, layers/resnet.py:172-201,
                   layers/attention.py:100-188, layers/embedding.py:52-105
  dit_forward      score_networks/dit.py:42-74,187-229, layers/attention.py:350-380,
                   layers/mlp.py:40-48, layers/embedding.py:325-406,455-504,
                   layers/utils.py:90-121
  pixart_forward   score_networks/pixart.py:76-120,227-268, layers/attention.py:209-228,
                   layers/embedding.py:202-237
  effunet_forward  score_networks/efficient_unet.py:219-256, layers/resnet.py:204-437,
                   layers/super_resolution.py:47-157
  unet3d_forward   score_networks/unet_3d.py:316-353, layers/resnet_3d.py:103-254,
                   layers/attention.py:457-676, layers/embedding.py:108-143
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- embeddings
def sinusoid_unet(t, dim, max_time):
    """[sin|cos], freq_i = exp(-ln(1e4) * i/(half-1)), arg = t*1000/max_time
    (layers/embedding.py:66-76)."""
    x = t * 1000.0 / max_time
    half = dim // 2
    f = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    a = x[:, None] * f[None, :]
    return torch.cat((a.sin(), a.cos()), dim=-1)


def sinusoid_dit(t, dim, max_period=10000):
    """[cos|sin], freq_i = exp(-ln(1e4) * i/half) (layers/utils.py:102-117)."""
    half = dim // 2
    f = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    a = t[:, None].float() * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


def sincos_pos_embed_2d(dim, grid, base_size=16, lewei_scale=1.0):
    """Fixed 2-D sin-cos position table (layers/utils.py:188-258): per token
    [sin(w*om) | cos(w*om) | sin(h*om) | cos(h*om)], om_k = 10000^(-k/(dim/4)),
    positions = arange(grid)/(grid/base_size)/lewei_scale, tokens row-major."""
    pos = np.arange(grid, dtype=np.float32) / (grid / base_size) / lewei_scale
    quarter = dim // 4
    om = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)
    ww, hh = np.meshgrid(pos, pos)           # ww[r, c] = pos[c], hh[r, c] = pos[r]

    def enc(p):
        o = np.einsum("m,d->md", p.reshape(-1), om)
        return np.concatenate([np.sin(o), np.cos(o)], axis=1)

    return torch.from_numpy(np.concatenate([enc(ww), enc(hh)], axis=1)).float()


def _lin(sd, pre, x, bias=True):
    return F.linear(x, sd[pre + ".weight"], sd[pre + ".bias"] if bias else None)


def _gn(sd, pre, x):
    return F.group_norm(x, 32, sd[pre + ".weight"], sd[pre + ".bias"], 1e-5)


# ----------------------------------------------------------------------------- UNet (2-D)
def unet_layout(p):
    """Replays the constructor's block numbering (score_networks/unet.py:139-245).
    Returns (downs, ups): lists of entries; each entry is a list of
    ("res", cin, cout) | ("attn", c) | ("down",) | ("up",)."""
    nf = p["num_features"]
    mults = p["channel_multipliers"]
    nres = p["num_resnet_blocks"]
    nres = nres if isinstance(nres, list) else [nres] * len(mults)
    size = p["input_spatial_size"]
    size = size[1] if isinstance(size, list) else size
    if "attention" in p:
        att_res = p["attention"]["attention_resolutions"]
    else:
        att_res = p["attention_resolutions"]
    att_ds = [size // int(r) for r in att_res]
    chans = [nf]
    ch, ds = nf, 1
    downs, ups = [], []
    for level, m in enumerate(mults):
        for _ in range(nres[level]):
            e = [("res", ch, m * nf)]
            ch = m * nf
            if ds in att_ds:
                e.append(("attn", ch))
            downs.append(e)
            chans.append(ch)
        if level != len(mults) - 1:
            downs.append([("down",)])
            chans.append(ch)
            ds *= 2
    mid_ch = ch
    for level, m in list(enumerate(mults))[::-1]:
        for i in range(nres[level] + 1):
            e = [("res", ch + chans.pop(), nf * m)]
            ch = nf * m
            if ds in att_ds:
                e.append(("attn", ch))
            if level and i == nres[level]:
                e.append(("up",))
                ds //= 2
            ups.append(e)
    return downs, mid_ch, ups


def _resblock2d(sd, pre, x, emb):
    h = F.conv2d(F.silu(_gn(sd, pre + "in_layers.0", x)), sd[pre + "in_layers.2.weight"],
                 sd[pre + "in_layers.2.bias"], padding=1)
    e = _lin(sd, pre + "emb_layers.1", F.silu(emb))
    scale, shift = e[:, :, None, None].chunk(2, dim=1)
    h = _gn(sd, pre + "out_layers.0", h) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[pre + "out_layers.3.weight"], sd[pre + "out_layers.3.bias"], padding=1)
    if pre + "skip_connection.weight" in sd:
        x = F.conv2d(x, sd[pre + "skip_connection.weight"], sd[pre + "skip_connection.bias"])
    return x + h


def qkv_attention_interleaved(qkv, heads):
    """QKVAttention (layers/attention.py:152-188): channel order [h0: q k v | h1: q k v ...],
    scale ch^-1/4 on both q and k, fp32 softmax."""
    bs, width, length = qkv.shape
    ch = width // (3 * heads)
    q, k, v = qkv.reshape(bs * heads, ch * 3, length).split(ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
    return torch.einsum("bts,bcs->bct", w, v).reshape(bs, -1, length)


def chan_layer_norm(x, g, dim):
    """layers/attention.py:284-310 (LayerNorm with an axis; gain only, eps 1e-5 in fp32, biased variance)."""
    var = torch.var(x, dim=dim, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=dim, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def _spatial_attn(sd, pre, x, dim_head=64, text=None):
    """SpatialCrossAttention (layers/attention.py:100-141).  ``text`` (B, L, Ctx): TextEmbeddingsAdapter with
    swap_context_channels (context.py:115-140) -> (B, Ctx, L), ChanLayerNorm over the channel axis, ``_encoder_kv``
    conv1d -> per-head [k | v], concatenated IN FRONT of the self-attention keys / values (attention.py:166-180)."""
    b, c = x.shape[:2]
    heads = c // dim_head
    qkv = F.conv1d(_gn(sd, pre + "_norm", x).view(b, c, -1), sd[pre + "_qkv.weight"], sd[pre + "_qkv.bias"])
    if text is not None and pre + "_encoder_kv.weight" in sd:
        ctx = text.permute(0, 2, 1)
        if pre + "_context_layer_norm.g" in sd:
            ctx = chan_layer_norm(ctx, sd[pre + "_context_layer_norm.g"], -2)
        ekv = F.conv1d(ctx, sd[pre + "_encoder_kv.weight"], sd[pre + "_encoder_kv.bias"])
        bs, width, length = qkv.shape
        ch = width // (3 * heads)
        q, k, v = qkv.reshape(bs * heads, ch * 3, length).split(ch, dim=1)
        ek, ev = ekv.reshape(bs * heads, ch * 2, -1).split(ch, dim=1)
        k, v = torch.cat([ek, k], dim=-1), torch.cat([ev, v], dim=-1)
        s = 1 / math.sqrt(math.sqrt(ch))
        w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
        a = torch.einsum("bts,bcs->bct", w, v).reshape(bs, -1, length)
    else:
        a = qkv_attention_interleaved(qkv, heads)
    h = F.conv1d(a, sd[pre + "_proj_out.weight"], sd[pre + "_proj_out.bias"])
    return x + h.reshape(x.shape)


def attention_pooling(sd, pre, x, heads):
    """AttentionPooling (layers/attention.py:231-283): a class token (mean over the sequence + positional embedding)
    attends over [class token, sequence]; q and k are both scaled by d^-1/4; returns the class token (B, C)."""
    bs, length, width = x.shape
    d = width // heads

    def shape(t):
        return t.view(bs, -1, heads, d).transpose(1, 2).reshape(bs * heads, -1, d).transpose(1, 2)

    cls = x.mean(dim=1, keepdim=True) + sd[pre + "positional_embedding"]
    xx = torch.cat([cls, x], dim=1)
    q, k, v = shape(_lin(sd, pre + "q_proj", cls)), shape(_lin(sd, pre + "k_proj", xx)), shape(_lin(sd, pre + "v_proj", xx))
    s = 1 / math.sqrt(math.sqrt(d))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v)
    return a.reshape(bs, -1, 1).transpose(1, 2)[:, 0, :]


def pooled_text_to_timestep(sd, p, text):
    """PooledTextEmbeddingsToTimestep (layers/embedding.py:146-169): LayerNorm -> AttentionPooling -> Linear -> LayerNorm,
    added to the timestep embedding.  Returns the (B, time_embedding_dim) addend, or None if the config has no such head."""
    head = p["conditioning"]["context_transformer_head"]
    head = head if isinstance(head, list) else [head]
    for n, h in enumerate(head):
        if h["target"].endswith("PooledTextEmbeddingsToTimestep"):
            pre = f"_context_transformers.{n}._encoder_pooling."
            y = F.layer_norm(text, text.shape[-1:], sd[pre + "0.weight"], sd[pre + "0.bias"], 1e-5)
            y = attention_pooling(sd, pre + "1.", y, h["params"]["attention_pooling_heads"])
            y = _lin(sd, pre + "2", y)
            return F.layer_norm(y, y.shape[-1:], sd[pre + "3.weight"], sd[pre + "3.bias"], 1e-5)
    return None


def time_input_key(p):
    """Which context entry feeds the timestep projection ("timestep" or "logsnr_t")."""
    head = p["conditioning"]["context_transformer_head"]
    head = head if isinstance(head, list) else [head]
    for h in head:
        if h["params"].get("projection_key") == "timestep":
            return h["params"]["input_context_key"]
    return "timestep"


def unet_time_embedding(sd, p, t):
    proj = p["conditioning"]["projections"]["timestep"]
    if proj["target"].endswith("InvCosTimestepEmbeddingProjection"):
        return invcos_time_embedding(sd, p, t)
    tp = proj["params"]
    s = sinusoid_unet(t, tp["num_features"], tp.get("max_time", 1000.0))
    pre = "_projections.timestep._projection."
    return _lin(sd, pre + "3", F.silu(_lin(sd, pre + "1", s)))


def unet_forward(sd, p, x, t, taps=None, text=None):
    """x (B,C,H,W) fp32; t (B,) = context[time_input_key(p)]: int64 loop index (discrete),
    fp32 time (flow) or fp32 logsnr_t (continuous).  ``taps`` (optional dict) receives
    named intermediate activations for layer-level parity tests.  ``text`` (B, L, Ctx): text embeddings of the
    text-conditioned configs (Imagen base / GLIDE: unet.py:87-98 with context_dim > 0)."""
    emb = unet_time_embedding(sd, p, t)
    if text is not None:
        pooled = pooled_text_to_timestep(sd, p, text)
        if pooled is not None:
            emb = emb + pooled
    downs, _, ups = unet_layout(p)
    h = F.conv2d(x, sd["_initial_convolution.weight"], None, padding=1)
    if taps is not None:
        taps["temb"] = emb
        taps["conv_in"] = h
    hs = [h]

    def run(entry, pre, h):
        for j, op in enumerate(entry):
            if op[0] == "res":
                h = _resblock2d(sd, f"{pre}.{j}.", h, emb)
            elif op[0] == "attn":
                h = _spatial_attn(sd, f"{pre}.{j}.", h, text=text)
            elif op[0] == "down":
                h = F.avg_pool2d(h, 2, 2)
            elif op[0] == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
        return h

    for n, entry in enumerate(downs):
        h = run(entry, f"downs.{n}", h)
        hs.append(h)
        if taps is not None:
            taps[f"downs.{n}"] = h
    h = _resblock2d(sd, "middle.0.", h, emb)
    h = _spatial_attn(sd, "middle.1.", h, text=text)
    h = _resblock2d(sd, "middle.2.", h, emb)
    if taps is not None:
        taps["middle"] = h
    for n, entry in enumerate(ups):
        h = run(entry, f"ups.{n}", torch.cat([h, hs.pop()], dim=1))
        if taps is not None:
            taps[f"ups.{n}"] = h
    h = F.silu(_gn(sd, "final_projection.0", h))
    return F.conv2d(h, sd["final_projection.2.weight"], None, padding=1)


# ----------------------------------------------------------------------------- efficient UNet (Imagen SR stage)
def _resblock_efficient(sd, pre, x):
    """ResnetBlockEfficient (layers/resnet.py:204-250): no time embedding inside, always a 1x1 skip, sum scaled by 0.7071."""
    p = pre + "_resnet_path."
    h = F.conv2d(F.silu(_gn(sd, p + "0", x)), sd[p + "2.weight"], sd[p + "2.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(sd, p + "3", h)), sd[p + "6.weight"], sd[p + "6.bias"], padding=1)
    r = F.conv2d(x, sd[pre + "_skip_connection.weight"], sd[pre + "_skip_connection.bias"]) + h
    return r * 0.7071


def effunet_forward(sd, p, x, t, text=None, augmentation_timestep=None):
    """score_networks/efficient_unet.py:219-256 with DBlock / UBlock of layers/resnet.py:253-437: every DBlock starts with a
    stride-2 conv, adds Linear(SiLU(temb)) per channel, runs its ResnetBlockEfficient chain and (where configured) a
    SpatialCrossAttention; every UBlock does the same and ends with nearest x2 + conv.  The timestep embedding carries the
    Gaussian-conditioning-augmentation level (layers/super_resolution.py:124-157)."""
    emb = unet_time_embedding(sd, p, t)
    head = p["conditioning"]["context_transformer_head"]
    head = head if isinstance(head, list) else [head]
    for n, hcfg in enumerate(head):
        if hcfg["target"].endswith("GaussianConditioningAugmentationToTimestep"):
            tp = hcfg["params"]
            pre = f"_context_transformers.{n}._embedding_projection._projection."
            s_emb = sinusoid_unet(augmentation_timestep, tp["num_features"], 1000.0)
            emb = emb + _lin(sd, pre + "3", F.silu(_lin(sd, pre + "1", s_emb)))
        elif hcfg["target"].endswith("PooledTextEmbeddingsToTimestep") and text is not None:
            emb = emb + pooled_text_to_timestep(sd, p, text)
    mults = p["channel_multipliers"]
    nres = p["num_resnet_blocks"]
    nres = nres if isinstance(nres, list) else [nres] * len(mults)
    h = F.conv2d(x, sd["_initial_convolution.weight"], None, padding=1)

    def block(pre, h, n_res):
        e = _lin(sd, pre + "_embedding_layers.1", F.silu(emb))
        h = h + e[:, :, None, None]
        for r in range(n_res):
            h = _resblock_efficient(sd, f"{pre}_resnet_blocks.{r}.", h)
        if pre + "_attention._qkv.weight" in sd:
            h = _spatial_attn(sd, pre + "_attention.", h, text=text)
        return h

    hs = []
    for lvl in range(len(mults)):
        pre = f"downs.{lvl}."
        h = F.conv2d(h, sd[pre + "_downsampling_convolution.weight"], sd[pre + "_downsampling_convolution.bias"],
                     stride=2, padding=1)
        h = block(pre, h, nres[lvl])
        hs.append(h)
    hs.pop()
    for idx, lvl in enumerate(reversed(range(len(mults)))):
        pre = f"ups.{idx}."
        h = h if idx == 0 else torch.cat([h, hs.pop()], dim=1)
        h = block(pre, h, nres[lvl] + 1)
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        h = F.conv2d(h, sd[pre + "_upsample.conv.weight"], sd[pre + "_upsample.conv.bias"], padding=1)
    h = F.silu(_gn(sd, "final_projection.0", h))
    return F.conv2d(h, sd["final_projection.2.weight"], None, padding=1)


def sr_input(x, low_resolution_images, size, tables, s, z_cond):
    """InputPreprocessor (layers/super_resolution.py:47-121): bilinear (antialias) resize of the [0, 1] low-resolution images,
    normalise to [-1, 1], q_sample at the augmentation timestep s with noise z_cond (scheduler.py:289-308), concatenate behind
    x on the channel axis."""
    low = F.interpolate(low_resolution_images, size=(size, size), mode="bilinear", antialias=True, align_corners=False)
    low = low * 2 - 1
    a = tables["sqrt_alphas_cumprod"][s][:, None, None, None]
    c = tables["sqrt_one_minus_alphas_cumprod"][s][:, None, None, None]
    return torch.cat([x, a * low + c * z_cond], dim=1)


# ----------------------------------------------------------------------------- EDM: DDPM++ (SongUNet)
def _edm_gn(sd, pre, x):
    """layers/edm.py GroupNorm: min(32, C // 4) groups, eps 1e-6 in SongUNet blocks (score_networks/edm.py:68)."""
    C = x.shape[1]
    return F.group_norm(x, min(32, C // 4), sd[pre + ".weight"], sd[pre + ".bias"], 1e-6)


def _edm_block(sd, pre, x, emb, skip_scale, up=False, down=False):
    """layers/edm.py UNetBlock.forward (:302-343) with adaptive_scale = False, resample_filter [1, 1] (up = nearest x2,
    down = 2x2 average: Conv2d.forward :97-145 with f = [[1,1],[1,1]] / 4), resample_proj = True, one attention head."""
    def resample(t):
        if up:
            return F.interpolate(t, scale_factor=2, mode="nearest")
        if down:
            return F.avg_pool2d(t, 2, 2)
        return t
    orig = x
    x = F.conv2d(resample(F.silu(_edm_gn(sd, pre + "norm0", x))), sd[pre + "conv0.weight"], sd[pre + "conv0.bias"], padding=1)
    params = _lin(sd, pre + "affine", emb)[:, :, None, None]
    x = F.silu(_edm_gn(sd, pre + "norm1", x + params))
    x = F.conv2d(x, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    if pre + "skip.weight" in sd:
        orig = F.conv2d(resample(orig), sd[pre + "skip.weight"], sd[pre + "skip.bias"])
    x = (x + orig) * skip_scale
    if pre + "qkv.weight" in sd:
        B, C = x.shape[:2]
        qkv = F.conv2d(_edm_gn(sd, pre + "norm2", x), sd[pre + "qkv.weight"], sd[pre + "qkv.bias"])
        q, k, v = qkv.reshape(B, C, 3, -1).unbind(2)                        # one head: channel c of q = output channel 3c
        w = torch.einsum("ncq,nck->nqk", q, k / math.sqrt(C)).softmax(dim=2)
        a = torch.einsum("nqk,nck->ncq", w, v)
        x = (F.conv2d(a.reshape(x.shape), sd[pre + "proj.weight"], sd[pre + "proj.bias"]) + x) * skip_scale
    return x


def songunet_forward(sd, p, x, noise_labels, prefix="model."):
    """score_networks/edm.py SongUNet.forward (:183-238), DDPM++ configuration (positional embedding, standard encoder /
    decoder).  ``noise_labels`` (n,) fp32 with n = 1 or B."""
    nc = p["model_channels"] * p.get("channel_mult_noise", 1)
    freqs = torch.arange(0, nc // 2, dtype=torch.float32) / (nc // 2 - 1)
    freqs = (1 / 10000) ** freqs
    e = noise_labels.float().ger(freqs)
    emb = torch.cat([e.cos(), e.sin()], dim=1)
    emb = emb.reshape(emb.shape[0], 2, -1).flip(1).reshape(*emb.shape)              # swap sin / cos
    emb = F.silu(_lin(sd, prefix + "map_layer0", emb))
    emb = F.silu(_lin(sd, prefix + "map_layer1", emb))
    s = math.sqrt(0.5)
    names = {"enc": [], "dec": []}
    for k in sd:                                                                    # ModuleDict order = state-dict order
        if k.startswith(prefix + "enc.") or k.startswith(prefix + "dec."):
            part, name = k[len(prefix):].split(".")[:2]
            if name not in names[part]:
                names[part].append(name)
    skips = []
    for name in names["enc"]:
        pre = f"{prefix}enc.{name}."
        if name.endswith("_conv"):
            x = F.conv2d(x, sd[pre + "weight"], sd[pre + "bias"], padding=1)
        else:
            x = _edm_block(sd, pre, x, emb, s, down=name.endswith("_down"))
        skips.append(x)
    out = None
    for name in names["dec"]:
        pre = f"{prefix}dec.{name}."
        if name.endswith("aux_norm"):
            out = _edm_gn(sd, pre[:-1], x)
        elif name.endswith("aux_conv"):
            out = F.conv2d(F.silu(out), sd[pre + "weight"], sd[pre + "bias"], padding=1)
        else:
            cin = sd[pre + "norm0.weight"].shape[0]
            if x.shape[1] != cin:
                x = torch.cat([x, skips.pop()], dim=1)
            x = _edm_block(sd, pre, x, emb, s, up=name.endswith("_up"))
    return out


# ----------------------------------------------------------------------------- DiT
def _ln(x):
    return F.layer_norm(x, x.shape[-1:], None, None, 1e-6)


def mhsa(sd, pre, x, heads):
    """MultiHeadSelfAttention, unfused branch (layers/attention.py:350-380)."""
    B, N, C = x.shape
    d = C // heads
    qkv = _lin(sd, pre + "qkv", x).reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    a = ((q * d ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    return _lin(sd, pre + "proj", (a @ v).transpose(1, 2).reshape(B, N, C))


def _mlp(sd, pre, x):
    return _lin(sd, pre + "fc2", F.gelu(_lin(sd, pre + "fc1", x), approximate="tanh"))


def patch_embed(sd, x, patch):
    y = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=patch)
    return y.flatten(2).transpose(1, 2)


def unpatchify(y, patch, c):
    n, T, _ = y.shape
    g = int(T ** 0.5)
    y = y.reshape(n, g, g, patch, patch, c)
    return torch.einsum("nhwpqc->nchpwq", y).reshape(n, c, g * patch, g * patch)


def dit_conditioning(sd, p, t, classes):
    fs = p["conditioning"]["projections"]["timestep"]["params"]["frequency_embedding_size"]
    pre = "_projections.timestep.mlp."
    temb = _lin(sd, pre + "2", F.silu(_lin(sd, pre + "0", sinusoid_dit(t, fs))))
    cemb = sd["_projections.classes.embedding_table.weight"][classes]
    return cemb + temb            # DiTCombineEmbeddngs: class_emb += t_emb


def dit_forward(sd, p, x, t, classes, taps=None):
    heads, patch = p["num_heads"], p["patch_size"]
    c = dit_conditioning(sd, p, t, classes)
    h = patch_embed(sd, x, patch) + sd["pos_embed"]
    if taps is not None:
        taps["c"] = c
        taps["tokens"] = h
    for n in range(p["depth"]):
        pre = f"blocks.{n}."
        s1, sc1, g1, s2, sc2, g2 = _lin(sd, pre + "adaLN_modulation.1", F.silu(c)).chunk(6, dim=1)
        h = h + g1[:, None] * mhsa(sd, pre + "attn.", _ln(h) * (1 + sc1[:, None]) + s1[:, None], heads)
        h = h + g2[:, None] * _mlp(sd, pre + "mlp.", _ln(h) * (1 + sc2[:, None]) + s2[:, None])
        if taps is not None:
            taps[f"block{n}"] = h
    sh, sc = _lin(sd, "final_layer.adaLN_modulation.1", F.silu(c)).chunk(2, dim=1)
    y = _lin(sd, "final_layer.linear", _ln(h) * (1 + sc[:, None]) + sh[:, None])
    return unpatchify(y, patch, p["input_channels"])


# ----------------------------------------------------------------------------- PixArt-alpha
def cross_attention(sd, pre, x, y, heads):
    """LastChannelCrossAttention (layers/attention.py:209-228): no k/v/q bias, no mask."""
    B, N, C = x.shape
    d = C // heads
    q = F.linear(x, sd[pre + "to_q.weight"]).view(B, N, heads, d).transpose(1, 2)
    k = F.linear(y, sd[pre + "to_k.weight"]).view(B, -1, heads, d).transpose(1, 2)
    v = F.linear(y, sd[pre + "to_v.weight"]).view(B, -1, heads, d).transpose(1, 2)
    a = (q @ k.transpose(-2, -1) * d ** -0.5).softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(sd, pre + "to_out", o)


def context_projection_key(sd):
    for k in sd:
        if k.endswith("y_proj.fc1.weight"):
            return k[: -len("fc1.weight")]
    raise KeyError("y_proj")


def pixart_forward(sd, p, x, t, text_embeddings, taps=None):
    heads, patch = p["num_heads"], p["patch_size"]
    fs = p["conditioning"]["projections"]["timestep"]["params"]["frequency_embedding_size"]
    pre = "_projections.timestep.mlp."
    temb = _lin(sd, pre + "2", F.silu(_lin(sd, pre + "0", sinusoid_dit(t, fs))))
    y = _mlp(sd, context_projection_key(sd), text_embeddings)
    h = patch_embed(sd, x, patch) + sd["pos_embed"]
    t0 = _lin(sd, "t_block.1", F.silu(temb))
    B = x.shape[0]
    if taps is not None:
        taps["temb"], taps["y"], taps["tokens"] = temb, y, h
    for n in range(p["depth"]):
        pre = f"blocks.{n}."
        s1, sc1, g1, s2, sc2, g2 = (sd[pre + "scale_shift_table"][None] + t0.reshape(B, 6, -1)).chunk(6, dim=1)
        h = h + g1 * mhsa(sd, pre + "attn.", _ln(h) * (1 + sc1) + s1, heads)
        h = h + cross_attention(sd, pre + "cross_attn.", h, y, heads)
        h = h + g2 * _mlp(sd, pre + "mlp.", _ln(h) * (1 + sc2) + s2)
        if taps is not None:
            taps[f"block{n}"] = h
    sh, sc = (sd["final_layer.scale_shift_table"][None] + temb[:, None]).chunk(2, dim=1)
    out = _lin(sd, "final_layer.linear", _ln(h) * (1 + sc) + sh)
    return unpatchify(out, patch, p["input_channels"])


# ----------------------------------------------------------------------------- UNet-3D (video)
def invcos_time_embedding(sd, p, logsnr):
    """InvCosTimestepEmbeddingProjection (layers/embedding.py:108-143)."""
    tp = p["conditioning"]["projections"]["timestep"]["params"]
    u = torch.arctan(torch.exp(-0.5 * torch.clip(logsnr, tp["clip_min"], tp["clip_max"]))) / (0.5 * np.pi)
    s = sinusoid_unet(u, tp["num_features"], tp["max_time"])
    pre = "_projections.timestep._projection."
    return _lin(sd, pre + "3", F.silu(_lin(sd, pre + "1", s)))


def _resblock3d(sd, pre, x, emb, mlp_layers):
    def conv(h, key):
        return F.conv3d(h, sd[key + ".weight"], sd[key + ".bias"], padding=(0, 1, 1))

    h = conv(F.silu(_gn(sd, pre + "in_layers.0", x)), pre + "in_layers.2")
    e = emb
    for j in range(mlp_layers):
        e = _lin(sd, f"{pre}emb_layers.{j}.fc2", F.silu(_lin(sd, f"{pre}emb_layers.{j}.fc1", e)))
    scale, shift = e[:, :, None, None, None].chunk(2, dim=1)
    h = _gn(sd, pre + "out_layers.0", h) * (1 + scale) + shift
    h = conv(F.silu(h), pre + "out_layers.3")
    if pre + "skip_connection.weight" in sd:
        x = F.conv3d(x, sd[pre + "skip_connection.weight"], sd[pre + "skip_connection.bias"])
    return x + h


def relpos_attention(qkv, heads, ek):
    """QKVAttentionWithRelativePosition (layers/attention.py:551-676): split
    (B,H,3,D,L); logits = q.k + q.E_k[h, j-i+L-1] with NO 1/sqrt(D); the (B,H,L,D)
    result is returned through a raw reshape(B,-1,L) with no transpose back."""
    B, C3, L = qkv.shape
    D = C3 // (3 * heads)
    q, k, v = (t.permute(0, 1, 3, 2) for t in qkv.reshape(B, heads, 3 * D, L).split(D, dim=2))
    logits = q @ k.transpose(-2, -1)
    M = ek.shape[1]                      # 2*max_rel-1
    rel = torch.einsum("bhld,hmd->bhlm", q, ek)       # (B,H,L,M)
    i = torch.arange(L)[:, None]
    j = torch.arange(L)[None, :]
    idx = (j - i + (M - 1) // 2).expand(B, heads, L, L)
    logits = logits + torch.gather(rel, 3, idx)
    a = torch.softmax(logits, dim=-1) @ v
    return a.reshape(B, -1, L)


def _temporal_attn(sd, pre, x, dim_head=64):
    """x (B,C,F,H,W) -> '(b h w) c f' -> GN over (C/32, F) -> qkv -> rel-pos attention."""
    B, C, Fr, H, W = x.shape
    y = x.permute(0, 3, 4, 1, 2).reshape(B * H * W, C, Fr)
    qkv = F.conv1d(_gn(sd, pre + "_norm", y), sd[pre + "_qkv.weight"], sd[pre + "_qkv.bias"])
    a = relpos_attention(qkv, C // dim_head, sd[pre + "_attention._k_embeddings_table"])
    h = F.conv1d(a, sd[pre + "_proj_out.weight"], sd[pre + "_proj_out.bias"])
    y = y + h
    return y.reshape(B, H, W, C, Fr).permute(0, 3, 4, 1, 2)


def _spatial_attn_video(sd, pre, x):
    B, C, Fr, H, W = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W)
    y = _spatial_attn(sd, pre, y)
    return y.reshape(B, Fr, C, H, W).permute(0, 2, 1, 3, 4)


def unet3d_forward(sd, p, x, logsnr_t, taps=None):
    """x (B,C,F,H,W); logsnr_t (B,) fp32 = context["logsnr_t"]."""
    emb = invcos_time_embedding(sd, p, logsnr_t)
    downs, _, ups = unet_layout(p)
    ml = p["mlp_layers"]
    h = F.conv3d(x, sd["_initial_convolution.weight"], None, padding=(0, 1, 1))
    hs = [h]

    def run(entry, pre, h):
        j = 0
        for op in entry:
            if op[0] == "res":
                h = _resblock3d(sd, f"{pre}.{j}.", h, emb, ml)
                j += 1
            elif op[0] == "attn":
                h = _spatial_attn_video(sd, f"{pre}.{j}.fn.", h)
                h = _temporal_attn(sd, f"{pre}.{j + 1}.fn.", h)
                j += 2
            elif op[0] == "down":
                h = F.avg_pool3d(h, (1, 2, 2), (1, 2, 2))
                j += 1
            elif op[0] == "up":
                h = F.interpolate(h, (h.shape[2], h.shape[3] * 2, h.shape[4] * 2), mode="nearest")
                j += 1
        return h

    for n, entry in enumerate(downs):
        h = run(entry, f"downs.{n}", h)
        hs.append(h)
        if taps is not None:
            taps[f"downs.{n}"] = h
    h = _resblock3d(sd, "middle.0.", h, emb, ml)
    h = _spatial_attn_video(sd, "middle.1.fn.", h)
    h = _temporal_attn(sd, "middle.2.fn.", h)
    h = _resblock3d(sd, "middle.3.", h, emb, ml)
    if taps is not None:
        taps["middle"] = h
    for n, entry in enumerate(ups):
        h = run(entry, f"ups.{n}", torch.cat([h, hs.pop()], dim=1))
        if taps is not None:
            taps[f"ups.{n}"] = h
    h = F.silu(_gn(sd, "final_projection.0", h))
    return F.conv3d(h, sd["final_projection.2.weight"], None, padding=(0, 1, 1))

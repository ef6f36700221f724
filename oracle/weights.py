"""Deterministic synthetic weights shared by the reference, the oracle and the
CUDA path.  TEST INFRASTRUCTURE ONLY.

The reference zero-initialises every resblock's second conv, every attention
output projection and all DiT adaLN/final layers (layers/resnet.py:154-158,
layers/attention.py:97, score_networks/dit.py:169-178), which makes random-init
parity vacuous (the DiT output is exactly 0).  So parity runs overwrite *every*
learnable tensor with values drawn from a per-key seeded generator; the
manifest (key -> shape) is the only thing a fixture has to carry.
"""
import hashlib

import torch

_KEEP = ("pos_embed", "resample_filter")          # deterministic tables: keep the constructor's value


def _seed(key: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def synth_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(_seed(key, seed))
    shape = tuple(shape)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if "embeddings_table" in key:                 # temporal rel-pos tables (H, 2L-1, D)
        return r * shape[-1] ** -0.5
    if "embedding_table" in key:                  # class-label table
        return r * 0.5
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r * fan_in ** -0.5
    if key.endswith("bias"):
        return r * 0.1
    return 1.0 + 0.1 * r                          # norm gains


#: The video model's temporal attention uses UNSCALED logits q.k (reference layers/attention.py:647).  With
#: unit-gain qkv weights |logit| ~ 8 and the softmax is one-hot -- a regime no trained checkpoint is in and one
#: that turns the test into a measurement of argmax ties.  Those qkv projections get gain 0.35 (logit std ~ 1).
TEMPORAL_QKV_GAIN = 0.35


def synth_state_dict(manifest, seed: int = 0):
    """manifest: {key: shape}.  Keys listed in _KEEP are skipped."""
    out = {k: synth_tensor(k, s, seed) for k, s in manifest.items() if not any(t in k for t in _KEEP)}
    for k in list(out):
        if k.endswith("._qkv.weight") and k[: -len("_qkv.weight")] + "_attention._k_embeddings_table" in manifest:
            out[k] = out[k] * TEMPORAL_QKV_GAIN
    return out


def canonical_manifest(state_dict, prefix="_score_network."):
    """Learnable score-network tensors of a reference/product state dict, without the
    alias copies registered under ``_context_transformers.N._projections`` (same storage)."""
    out = {}
    for k, v in state_dict.items():
        if not k.startswith(prefix):
            continue
        kk = k[len(prefix):]
        if kk.startswith("_context_transformers.") and "._projections." in kk:
            continue
        out[kk] = tuple(v.shape)
    return out

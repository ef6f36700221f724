"""LoRA on-ramp for the sampling path: low-rank deltas are MERGED into the base weights.

The reference (xdiffusion/lora.py:228-343, used by sampling/image/mnist/sample.py:86-98) swaps every Linear / Conv1d /
Conv2d found under a ``ResnetBlockBigGAN`` / ``SpatialCrossAttention`` / ... ancestor for a ``LoraInjected*`` wrapper that
evaluates ``W x + scale * up(down(x))`` with two extra small convolutions per layer.  At sampling time (eval mode, dropout
off, identity selector) that sum is exactly one layer with

    Linear / Conv1d(k=1):  W' = W + scale * up @ down
    Conv2d(k x k):         W'[o, i, kh, kw] = W[o, i, kh, kw] + scale * sum_r up[o, r, 0, 0] * down[r, i, kh, kw]

so the B200 path adds the delta to the parameter IN PLACE: the tcgen05 kernels keep running one implicit-GEMM per layer,
the repacked bf16 weights and any captured sampling loop are invalidated through the parameters' version counters
(layers/utils.py:Packed, diffusion/ddpm.py:_weights_fingerprint).  A ``.pt`` LoRA file is the flat list
``[up_0, down_0, up_1, down_1, ...]`` written by the reference's ``save_lora_weights`` (lora.py:312-324), in the order of
its module traversal ``_find_modules`` (lora.py:379-423), which ``iter_lora_targets`` reproduces (same class names, same
sub-module registration order).
"""
from typing import Iterator, List, Optional, Set, Tuple

import torch

UNET_DEFAULT_TARGET_REPLACE = {"CrossAttention", "Attention", "GEGLU", "SpatialCrossAttention", "ResnetBlockBigGAN",
                               "ResnetBlockDDPM"}
DEFAULT_TARGET_REPLACE = UNET_DEFAULT_TARGET_REPLACE
_SEARCH = (torch.nn.Linear, torch.nn.Conv1d, torch.nn.Conv2d)


def iter_lora_targets(model: torch.nn.Module, target_replace_module: Set[str] = DEFAULT_TARGET_REPLACE
                      ) -> Iterator[Tuple[str, torch.nn.Module]]:
    """(qualified name, layer) of every Linear / Conv1d / Conv2d below a module whose class name is in
    ``target_replace_module``, in the reference's traversal order."""
    for aname, ancestor in model.named_modules():
        if ancestor.__class__.__name__ not in target_replace_module:
            continue
        for name, module in ancestor.named_modules():
            if module.__class__ in _SEARCH:
                yield (f"{aname}.{name}" if aname else name), module


def lora_shapes(model, r: int = 4, target_replace_module: Set[str] = DEFAULT_TARGET_REPLACE) -> List[Tuple[tuple, tuple]]:
    """[(up shape, down shape)] a rank-r LoRA file for this model holds (the shapes LoraInjected* would create)."""
    out = []
    for _, m in iter_lora_targets(model, target_replace_module):
        if isinstance(m, torch.nn.Linear):
            out.append(((m.out_features, r), (r, m.in_features)))
        else:
            k = tuple(m.kernel_size)
            out.append(((m.out_channels, r) + (1,) * len(k), (r, m.in_channels // m.groups) + k))
    return out


def lora_delta(up: torch.Tensor, down: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """scale * up o down as a weight of the base layer's shape."""
    r = down.shape[0]
    d = up.reshape(up.shape[0], r).float() @ down.reshape(r, -1).float()
    return (scale * d).reshape((up.shape[0],) + tuple(down.shape[1:]))


@torch.no_grad()
def merge_lora_weights(model: torch.nn.Module, loras: List[torch.Tensor], scale: float = 1.0,
                       target_replace_module: Set[str] = DEFAULT_TARGET_REPLACE) -> List[str]:
    """Add the deltas of ``loras`` = [up_0, down_0, up_1, down_1, ...] to the matching base weights, in place.  Returns the
    names of the merged layers; the deltas are remembered on the model so that ``remove_lora_weights`` can undo them."""
    loras = list(loras)
    targets = list(iter_lora_targets(model, target_replace_module))
    if len(loras) != 2 * len(targets):
        raise ValueError(f"LoRA file holds {len(loras)} tensors, the model has {len(targets)} target layers (x2)")
    applied = []
    for name, layer in targets:
        up, down = loras.pop(0), loras.pop(0)
        delta = lora_delta(up.detach(), down.detach(), scale)
        if tuple(delta.shape) != tuple(layer.weight.shape):
            raise ValueError(f"LoRA shapes {tuple(up.shape)} x {tuple(down.shape)} do not fit {name} {tuple(layer.weight.shape)}")
        delta = delta.to(device=layer.weight.device, dtype=layer.weight.dtype)
        layer.weight.add_(delta)                       # bumps _version: repacked weights + captured loops are rebuilt
        applied.append((layer, delta))
    model.__dict__.setdefault("_xdb_lora_applied", []).extend(applied)
    return [n for n, _ in targets]


def load_lora_weights(model: torch.nn.Module, lora_path: str, scale: float = 1.0) -> List[str]:
    """Drop-in for the reference's ``load_lora_weights(model, lora_path)`` (lora.py:327-328)."""
    loras = torch.load(lora_path, map_location="cpu", weights_only=False)
    return merge_lora_weights(model, loras, scale=scale)


@torch.no_grad()
def remove_lora_weights(model: torch.nn.Module) -> int:
    """Subtract every merged delta again (the base weights return to their values up to fp32 rounding)."""
    applied = model.__dict__.pop("_xdb_lora_applied", [])
    for layer, delta in applied:
        layer.weight.sub_(delta)
    return len(applied)


def synth_lora(shapes, seed: int = 0, gain: float = 1.0) -> List[torch.nn.Parameter]:
    """Seeded random LoRA tensors for tests and benchmarks (a trained ``up`` is not zero; the reference initialises it to
    zero, which would make every comparison vacuous)."""
    g = torch.Generator().manual_seed(seed)
    out: List[torch.nn.Parameter] = []
    for up_shape, down_shape in shapes:
        r = down_shape[0]
        fan = 1
        for s in down_shape[1:]:
            fan *= s
        out.append(torch.nn.Parameter(torch.randn(up_shape, generator=g) * gain * r ** -0.5))
        out.append(torch.nn.Parameter(torch.randn(down_shape, generator=g) * fan ** -0.5))
    return out

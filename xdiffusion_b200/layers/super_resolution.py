"""Low-resolution conditioning of a cascade's super-resolution stage (reference: layers/super_resolution.py:10-157).

``InputPreprocessor`` turns the previous stage's [0, 1] samples into the extra input channels of the score network: bilinear
(antialiased) resize to the stage's resolution, normalisation to [-1, 1] -- both timestep-invariant, done once per sampling loop
-- and the Gaussian conditioning augmentation ``q_sample(low, s)``, whose noise the reference RE-DRAWS at every network
evaluation; that part and the channel concatenation are one kernel per evaluation (``xd_sr_input``, injected noise table or
in-kernel Philox).  ``GaussianConditioningAugmentationToTimestep`` adds the embedding of the augmentation level to the timestep
embedding.
"""
from typing import Dict

import torch

from .embedding import TimestepEmbeddingProjection


class InputPreprocessor(torch.nn.Module):
    LOW_KEY = "_xdb_sr_low"            # resized + normalised low-resolution images in the loop's static conditioning
    AUG_KEY = "_xdb_sr_aug"            # (s tensor [B] int64, a, c) of the augmentation level

    def __init__(self, low_resolution_size: int, super_resolution_size: int, context_input_key: str,
                 apply_gaussian_conditioning_augmentation: bool, is_spatial: bool = True, is_temporal: bool = False,
                 **kwargs):
        super().__init__()
        if is_temporal or not is_spatial:
            raise NotImplementedError("temporal super-resolution")
        self._super_resolution_size, self._low_resolution_size = super_resolution_size, low_resolution_size
        self._context_input_key = context_input_key
        self._apply_gaussian_conditioning_augmentation = apply_gaussian_conditioning_augmentation

    # ---- timestep-invariant part -------------------------------------------------------------------------------------
    @torch.no_grad()
    def _resize(self, low):
        assert low.dim() == 4 and low.shape[2] == self._low_resolution_size and low.shape[3] == self._low_resolution_size
        s = self._super_resolution_size
        up = torch.nn.functional.interpolate(low.float(), size=(s, s), mode="bilinear", antialias=True, align_corners=False)
        return (up * 2 - 1).contiguous()

    def _augmentation(self, context, noise_scheduler, B, device):
        """(s [B] int64 on the device, a, c): q_sample coefficients at the sampling augmentation level
        (super_resolution.py:90-111, scheduler.py:289-308)."""
        if not self._apply_gaussian_conditioning_augmentation:
            return None, 1.0, 0.0
        if "augmentation_level" not in context:
            raise NotImplementedError("random augmentation timesteps (training behaviour) at sampling time")
        if noise_scheduler.continuous():
            raise NotImplementedError("Gaussian conditioning augmentation with a continuous scheduler")
        s = (torch.ones(B, dtype=torch.long) * noise_scheduler.steps() * context["augmentation_level"]).to(torch.long)
        i = int(s[0])
        a = float(noise_scheduler.sqrt_alphas_cumprod[i])
        c = float(noise_scheduler.sqrt_one_minus_alphas_cumprod[i])
        return s.to(device), a, c

    def precompute(self, context: Dict, noise_scheduler):
        """Called by the sampling loop on its own static conditioning dict (at construction and after every refresh)."""
        if self._context_input_key not in context:
            return
        low = self._resize(context[self._context_input_key])
        old = context.get(self.LOW_KEY)
        if old is not None and old.shape == low.shape:
            old.copy_(low)
        else:
            context[self.LOW_KEY] = low
        context[self.AUG_KEY] = self._augmentation(context, noise_scheduler, low.shape[0], low.device)

    # ---- per network evaluation ----------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, context: Dict, noise_scheduler, **kwargs):
        low = context.get(self.LOW_KEY)
        if low is None:
            low = self._resize(context[self._context_input_key].to(x.device))
        aug = context.get(self.AUG_KEY) or self._augmentation(context, noise_scheduler, x.shape[0], x.device)
        s, a, c = aug
        if s is not None:
            context["augmentation_timestep"] = s              # read by GaussianConditioningAugmentationToTimestep
        assert low.shape[0] == x.shape[0] and low.shape[2:] == x.shape[2:]
        out = torch.empty((x.shape[0], x.shape[1] + low.shape[1]) + tuple(x.shape[2:]), device=x.device, dtype=torch.float32)
        idx = context.get("timestep_idx", 0)
        idx_dev, idx_host = (idx, -1) if torch.is_tensor(idx) else (None, int(idx))
        z, stride = context.get("sr_noise"), 0
        if z is not None and z.dim() == low.dim() + 1:          # [N, B, ...] per-step table, row = loop index
            stride = low.numel()
        torch.ops.xdb200.sr_input(x.contiguous(), low, z, stride, out, a, c, idx_dev, idx_host, int(context.get("seed", 0)),
                                  context.get("seed_dev"), int(context.get("row_offset", 0)) * low[0].numel())
        return out


class GaussianConditioningAugmentationToTimestep(torch.nn.Module):
    """timestep_embedding += TimestepEmbeddingProjection(augmentation_timestep)  (super_resolution.py:124-157)."""

    def __init__(self, num_features: int, time_embedding_mult: int, **kwargs):
        super().__init__()
        self._embedding_projection = TimestepEmbeddingProjection(num_features, time_embedding_mult)

    def forward(self, context: Dict, **kwargs):
        assert "timestep_embedding" in context and "augmentation_timestep" in context
        projection = self._embedding_projection(context["augmentation_timestep"]).contiguous()
        t = context["timestep_embedding"].contiguous()
        out = torch.empty_like(t)
        torch.ops.xdb200.add_rows_periodic(t, projection, t.shape[0], out)
        context["timestep_embedding"] = out
        return context

"""Rectified-flow Euler step (reference: samplers/rectified_flow.py:16-85): x <- x + v * (1/N)."""
from typing import Dict, Optional

import torch

from .. import ops
from .base import ReverseProcessSampler


class AncestralSampler(ReverseProcessSampler):
    def __init__(self, **kwargs):
        super().__init__()

    @staticmethod
    def network_time(timestep_idx: int, N: int, T: float, eps: float = 1e-3) -> float:
        """python-double time of loop index i, exactly as rectified_flow.py:52-56."""
        k = N - (timestep_idx + 1)
        return k / N * (T - eps) + eps

    @torch.no_grad()
    def p_sample(self, x: torch.Tensor, context: Dict, unconditional_context: Optional[Dict], diffusion_model,
                 guidance_fn=None, classifier_free_guidance: Optional[float] = None):
        sde = diffusion_model.sde()
        assert sde is not None
        if sde.sigma_t(0.5) != 0.0:
            raise NotImplementedError("stochastic (sigma_t > 0) flow sampling")
        idx = context["timestep_idx"]
        if not torch.is_tensor(idx):
            # eager path: rebuild context["timestep"] like the reference; in the captured loop the
            # per-step table already holds these values (GaussianDiffusion_DDPM._time_tables).
            context["timestep"] = torch.ones(x.shape[0], device=x.device) * self.network_time(idx, sde.N, sde.T)
        o = diffusion_model.predict_score(x, context=context)
        return self._launch_step(ops.MODE_EULER, x, o, context, diffusion_model, "euler")

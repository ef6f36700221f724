"""DDPM ancestral sampling step (reference: samplers/ancestral.py:21-72,189-324) on the fused kernel."""
from typing import Dict, Optional

import torch

from .. import ops
from .base import ReverseProcessSampler


class AncestralSampler(ReverseProcessSampler):
    def __init__(self, reconstruction_guidance: bool = False, omega: float = 2.0, noise_source: str = "philox",
                 seed: int = 0, **kwargs):
        super().__init__()
        self._reconstruction_guidance = reconstruction_guidance
        self._reconstruction_omega = omega
        self.noise_source, self.seed = noise_source, seed

    @torch.no_grad()
    def p_sample(self, x: torch.Tensor, context: Dict, unconditional_context: Optional[Dict], diffusion_model,
                 guidance_fn=None, classifier_free_guidance: Optional[float] = None):
        if guidance_fn is not None:
            raise NotImplementedError("classifier guidance needs autograd through the network (out of scope)")
        if self._reconstruction_guidance and "x_a" in context:
            raise NotImplementedError("reconstruction guidance needs autograd through the network (out of scope)")
        assert context["timestep"].shape == (x.shape[0],)
        o = self._score(diffusion_model, x, context, unconditional_context, classifier_free_guidance)
        return self._launch_step(ops.MODE_ANCESTRAL, x, o, context, diffusion_model, "ancestral")

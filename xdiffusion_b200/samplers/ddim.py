"""DDIM step in logSNR form (reference: samplers/ddim.py:16-123); continuous scheduler only."""
from typing import Dict, Optional

import torch

from .. import ops
from .base import ReverseProcessSampler


class DDIMSampler(ReverseProcessSampler):
    def __init__(self, **kwargs):
        super().__init__()

    @torch.no_grad()
    def p_sample(self, x: torch.Tensor, context: Dict, unconditional_context: Optional[Dict], diffusion_model,
                 guidance_fn=None, classifier_free_guidance: Optional[float] = None, clip_denoised: bool = True):
        if not diffusion_model.noise_scheduler().continuous():
            raise KeyError("logsnr_t")          # the reference fails the same way on a discrete scheduler
        if not clip_denoised:
            raise NotImplementedError("clip_denoised=False")
        o = self._score(diffusion_model, x, context, unconditional_context, classifier_free_guidance)
        return self._launch_step(ops.MODE_DDIM, x, o, context, diffusion_model, "ddim")

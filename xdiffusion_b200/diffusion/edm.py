"""Config-driven EDM diffusion model: sampling only (reference: diffusion/edm.py:30-95)."""
from typing import Callable, Dict, List, Optional, Tuple

import torch

from ..utils import DotConfig, instantiate_from_config, unnormalize_to_zero_to_one


class GaussianDiffusion_EDM(torch.nn.Module):
    def __init__(self, config: DotConfig):
        super().__init__()
        self._config = config
        self._score_network = instantiate_from_config(config.diffusion.score_network.to_dict())
        self._sampler = instantiate_from_config(config.diffusion.sampling.to_dict())

    def config(self) -> DotConfig:
        return self._config

    def sample(self, context: Optional[Dict] = None, num_samples: int = 16, guidance_fn: Optional[Callable] = None,
               classifier_free_guidance: Optional[float] = None, sampler=None, num_sampling_steps: Optional[int] = None,
               initial_noise: Optional[torch.Tensor] = None, context_preprocessor=None, noise=None,
               ) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]]]:
        """Same contract as the reference (diffusion/edm.py:58-95): latents ~ N(0, I) -> sampler.p_sample_loop -> [0, 1].
        ``initial_noise`` / ``noise`` (per-step churn noise) are accepted for parity runs."""
        if guidance_fn is not None or classifier_free_guidance is not None:
            raise NotImplementedError("guidance is not part of the reference's EDM sampling path")
        s = self._config.diffusion.sampling
        shape = (num_samples, s.output_channels, s.output_spatial_size, s.output_spatial_size)
        if "output_frames" in s:
            shape = shape[:2] + (s.output_frames,) + shape[2:]
        device = next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("xdiffusion_b200 runs on CUDA (sm_100a) only; move the model with .to('cuda')")
        self.eval()
        x_t = initial_noise.to(device) if initial_noise is not None else torch.randn(shape, device=device)
        x_0 = (sampler or self._sampler).p_sample_loop(diffusion_model=self, latents=x_t, class_labels=None, noise=noise)
        self.train()
        return unnormalize_to_zero_to_one(x_0.float()), None

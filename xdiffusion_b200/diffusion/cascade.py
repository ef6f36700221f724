"""Cascade of diffusion stages: sampling only (reference: diffusion/cascade.py:17-179).

Each stage is a ``GaussianDiffusion_DDPM`` built from its own config; ``sample()`` chains them, handing the [0, 1] samples of a
stage to the next one under ``config.super_resolution.conditioning_key``.  Stage configs are given as in the reference by a YAML
path (``diffusion_cascade.cascade_layer_N.config``), or inline as a dict.
"""
from typing import Callable, Dict, List, Optional, Tuple

import torch

from ..utils import DotConfig, load_yaml
from .ddpm import GaussianDiffusion_DDPM


class GaussianDiffusionCascade(torch.nn.Module):
    def __init__(self, config: DotConfig):
        super().__init__()
        self._config = config
        self._layers = torch.nn.ModuleList()
        idx = 1
        while f"cascade_layer_{idx}" in config.diffusion_cascade:
            stage = config.diffusion_cascade[f"cascade_layer_{idx}"].config
            if isinstance(stage, str):
                stage = load_yaml(stage)
            elif not isinstance(stage, DotConfig):
                stage = DotConfig(stage)
            self._layers.append(GaussianDiffusion_DDPM(stage))
            idx += 1

    def models(self) -> List[GaussianDiffusion_DDPM]:
        return list(self._layers)

    def config(self) -> DotConfig:
        return self._config

    def load_checkpoint(self, checkpoint_path: str):
        assert False, "Loading model weights for a cascade not supported yet."      # as the reference (cascade.py:71-72)

    def sample(self, context: Optional[Dict] = None, num_samples: int = 16, guidance_fn: Optional[Callable] = None,
               classifier_free_guidance: Optional[float] = None, sampler=None, initial_noise: Optional[torch.Tensor] = None,
               context_preprocessor=None, stage_kwargs: Optional[List[Dict]] = None,
               ) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]]]:
        """reference: cascade.py:148-179.  ``stage_kwargs`` (one dict per stage: num_sampling_steps, initial_noise, noise,
        cond_noise, seed) is an addition for parity runs."""
        assert initial_noise is None
        assert sampler is None
        outputs, previous = [], None
        for n, model in enumerate(self.models()):
            ctx = context.copy() if context is not None else {}
            if previous is not None:
                ctx[model.config().super_resolution.conditioning_key] = previous
            previous, _ = model.sample(context=ctx, num_samples=num_samples, guidance_fn=guidance_fn,
                                       classifier_free_guidance=classifier_free_guidance,
                                       **(stage_kwargs[n] if stage_kwargs else {}))
            outputs.append(previous)
        return previous, outputs

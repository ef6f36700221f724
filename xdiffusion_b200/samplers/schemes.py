"""Frame-index schedules for long-video sampling and the autoregressive driver built on them.

Drop-in for ``xdiffusion.samplers.schemes`` (reference: samplers/schemes.py:5-126): a scheme is an iterator that yields,
per sampling stage, the absolute indices of the frames the model is conditioned on (already generated or observed), the
indices of the frames it generates, and the boolean temporal mask (True = generate, False = keep the conditioning frame)
that ``GaussianDiffusion_DDPM.sample()`` blends with after every reverse-process step (``context["video_mask"]`` /
``context["x0"]``, reference diffusion/ddpm.py:963-982 -- on this path one fused kernel, xd_blend_frames).
``sample_with_scheme`` is the loop of the reference's sampling script (sampling/video/moving_mnist/sample.py:108-188)
without its file output.
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


class SamplingSchemeBase:
    def __init__(self, video_length: int, num_observed_frames: int, max_frames: int, step_size: int):
        """video_length: frames of the final video; num_observed_frames: frames given at the start; max_frames: frames
        (observed + latent) the model takes per call; step_size: new frames per stage."""
        self._video_length = video_length
        self._max_frames = max_frames
        self._num_obs = num_observed_frames
        self._done_frames = set(range(self._num_obs))
        self._obs_frames = list(range(self._num_obs))
        self._step_size = step_size
        self._current_step = 0
        self.B = None

    def get_unconditional_indices(self) -> List[int]:
        return list(range(self._max_frames))

    def next_indices(self) -> Tuple[List[int], List[int]]:
        raise NotImplementedError

    def __iter__(self):
        self.step = 0
        return self

    def __next__(self):
        if self.is_done():
            raise StopIteration
        first_unconditional = self._num_obs == 0 and self._current_step == 0
        if first_unconditional:          # nothing observed yet: one full window of latent frames, then continue as usual
            obs, latent = [], self.get_unconditional_indices()
        else:
            obs, latent = self.next_indices()
        assert isinstance(obs, list) and isinstance(latent, list)
        for idx in obs:
            assert idx in self._done_frames, (
                f"Attempting to condition on frame {idx} while it is not generated yet.\n"
                f"Generated frames: {self._done_frames}\nObserving: {obs}\nGenerating: {latent}")
        assert np.all(np.array(latent) < self._video_length)
        self._done_frames.update(latent)
        if first_unconditional:
            self._obs_frames = latent
        self._current_step += 1
        nb = self.B if self.B is not None else None
        obs_b = [obs] * nb if nb is not None else obs
        latent_b = [latent] * nb if nb is not None else latent
        # True = latent (generate), False = observed; window-relative positions
        rows = obs_b if nb is not None else obs
        mask = torch.ones((len(rows), self._max_frames), dtype=torch.bool)
        offset = self._step_size * (self._current_step - 1)
        for bi in range(len(rows)):
            for frame in rows[bi]:
                rel = frame - offset
                assert 0 <= rel < self._max_frames
                mask[bi][rel] = False
        return obs_b, latent_b, mask

    def is_done(self) -> bool:
        return len(self._done_frames) >= self._video_length

    @property
    def typename(self):
        return type(self).__name__

    def set_videos(self, videos):
        self.B = len(videos)

    @property
    def num_observations(self):
        return self._num_obs

    @property
    def video_length(self):
        return self._video_length


class Autoregressive(SamplingSchemeBase):
    """Condition on the last (max_frames - step_size) finished frames, generate the next step_size."""

    def next_indices(self):
        if len(self._done_frames) == 0:
            return [], list(range(self._max_frames))
        obs = sorted(self._done_frames)[-(self._max_frames - self._step_size):]
        first = obs[-1] + 1
        return obs, list(range(first, min(first + self._step_size, self._video_length)))


@torch.no_grad()
def sample_with_scheme(diffusion_model, scheme: SamplingSchemeBase, num_samples: int, channels: int, image_size,
                       context: Optional[Dict] = None, **sample_kwargs) -> torch.Tensor:
    """Generate ``scheme.video_length`` frames stage by stage (reference sampling/video/moving_mnist/sample.py:108-188):
    every stage runs the full reverse process on a window of ``max_frames`` frames whose observed part is held fixed by the
    video-mask blend.  Returns [B, C, video_length, H, W] in [0, 1]."""
    assert scheme.num_observations == 0, "the reference driver starts unconditionally"
    device = next(diffusion_model.parameters()).device
    hw = list(image_size) if isinstance(image_size, (list, tuple)) else [image_size, image_size]
    context = dict(context or {})
    samples = torch.zeros((num_samples, scheme.video_length, channels, hw[0], hw[1]))      # frames-first, like the reference
    it = iter(scheme)
    while True:
        it.set_videos(samples)
        try:
            obs_idx, latent_idx, temporal_mask = next(it)
        except StopIteration:
            break
        frame_indices = torch.cat([torch.tensor(obs_idx), torch.tensor(latent_idx)], dim=1).long()
        x0 = torch.stack([samples[i, fi] for i, fi in enumerate(frame_indices)], dim=0).clone()
        context["x0"] = (x0.permute(0, 2, 1, 3, 4) * 2 - 1).to(device)                  # channels first, [-1, 1]
        context["frame_indices"] = frame_indices.to(device)
        context["video_mask"] = temporal_mask.to(device)
        local, _ = diffusion_model.sample(num_samples=num_samples, context=context, **sample_kwargs)
        local = local.permute(0, 2, 1, 3, 4).cpu()
        for i, li in enumerate(latent_idx):
            samples[i, li] = local[i, -len(li):]
    return samples.permute(0, 2, 1, 3, 4)

"""Noise schedulers of the sampling path (reference: xdiffusion/scheduler.py).

Same class names, constructor arguments and registered buffer names as the reference (the buffers
are part of its checkpoints, scheduler.py:185-224,396-399).  The tables are built once on the host
with the same fp64->fp32 recipe and are only *read* on the hot path: ``step_coefficients`` gathers
them into one row of eight fp32 coefficients per loop index for the fused sampler-step kernel
(csrc/step.cu), and ``network_time_tables`` provides the per-step network inputs.
"""
from typing import Dict, Tuple

import numpy as np
import torch

from .utils import instantiate_from_config


def _linear_betas(T, lo, hi):
    s = 1000 / T
    return torch.linspace(s * lo, s * hi, T, dtype=torch.float64)


def _cosine_betas(T, s=0.008):
    x = torch.linspace(0, T, T + 1, dtype=torch.float64)
    ac = torch.cos(((x / T) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def _quadratic_betas(T, lo, hi):
    s = 1000 / T
    return torch.linspace((s * lo) ** 0.5, (s * hi) ** 0.5, T, dtype=torch.float64) ** 2


def _sigmoid_betas(T, lo, hi):
    s = 1000 / T
    return torch.sigmoid(torch.linspace(-6, 6, T, dtype=torch.float64)) * (s * hi - s * lo) + s * lo


class NoiseScheduler(torch.nn.Module):
    """Interface kept from the reference (scheduler.py:69-124); training-time methods are out of scope."""

    def continuous(self) -> bool:
        raise NotImplementedError

    def steps(self) -> int:
        raise NotImplementedError

    def sample_random_times(self, batch_size, device):
        raise NotImplementedError("training-time API (out of scope for the sampling hot path)")

    q_sample = predict_v_from_x_and_epsilon = update_with_all_losses = sample_random_times


class DiscreteNoiseScheduler(NoiseScheduler):
    def __init__(self, schedule_type: str, num_scales: int, loss_type: str = "l2", min_beta: float = 0.0001,
                 max_beta: float = 0.02, importance_sampler: Dict = {}, **kwargs):
        super().__init__()
        T = num_scales
        if schedule_type == "linear":
            betas = _linear_betas(T, min_beta, max_beta)
        elif schedule_type == "cosine":
            betas = _cosine_betas(T)
        elif schedule_type == "quadratic":
            betas = _quadratic_betas(T, min_beta, max_beta)
        elif schedule_type == "sigmoid":
            betas = _sigmoid_betas(T, min_beta, max_beta)
        elif schedule_type == "jsd":
            betas = 1.0 / torch.linspace(T, 1, T)
        else:
            raise NotImplementedError(f"Noise schedule {schedule_type} not implemented.")
        if importance_sampler:
            self._importance_sampler = instantiate_from_config(importance_sampler)
        self.num_timesteps = int(betas.shape[0])
        self.loss_type = loss_type
        alphas = 1.0 - betas
        ac = torch.cumprod(alphas, 0)
        ac_prev = torch.nn.functional.pad(ac[:-1], (1, 0), value=1.0)
        pv = betas * (1.0 - ac_prev) / (1.0 - ac)
        buffers = {
            "betas": betas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
            "sqrt_alphas_cumprod": torch.sqrt(ac),
            "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
            "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
            "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
            "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
            "posterior_variance": pv,
            "posterior_log_variance_clipped": torch.log(pv.clamp(min=1e-20)),
            "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
            "posterior_mean_coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
        }
        for name, val in buffers.items():
            self.register_buffer(name, val.to(torch.float32))

    def steps(self) -> int:
        return self.num_timesteps

    def continuous(self) -> bool:
        return False

    def fixed_large_log_variance(self) -> torch.Tensor:
        """log(cat(posterior_variance[1], betas[1:])), fp32 on the fp32 buffers (scheduler.py:244-254)."""
        return torch.log(torch.cat([self.posterior_variance[1:2], self.betas[1:]]))

    def step_coefficients(self, prediction: str, num_sampling_steps: int, sampler: str = "ancestral"):
        """fp32 [T, 8] rows (a, b, c1, c2, sigma, e1, e2, -) and the x0 form for csrc/step.cu.
        Loop index i reads row i of the FULL table: the reference does not respace
        (diffusion/ddpm.py:919-934)."""
        if sampler != "ancestral":
            raise NotImplementedError("DDIM needs the continuous scheduler (reference: samplers/ddim.py:43-45)")
        # Host (CPU) torch ops on purpose: the same libm results as the reference's CPU path.
        b = {k: v.detach().cpu() for k, v in self.named_buffers()}
        c = torch.zeros(self.num_timesteps, 8, dtype=torch.float32)
        if prediction == "epsilon":
            c[:, 0], c[:, 1] = b["sqrt_recip_alphas_cumprod"], b["sqrt_recipm1_alphas_cumprod"]
        elif prediction == "v":
            c[:, 0], c[:, 1] = b["sqrt_alphas_cumprod"], b["sqrt_one_minus_alphas_cumprod"]
        else:
            raise NotImplementedError(prediction)
        c[:, 2], c[:, 3] = b["posterior_mean_coef1"], b["posterior_mean_coef2"]
        c[:, 4] = torch.exp(0.5 * torch.log(torch.cat([b["posterior_variance"][1:2], b["betas"][1:]])))
        return c, 0

    def network_time_tables(self, num_sampling_steps: int):
        """context["timestep"] per loop index: the int64 index itself (ddpm.py:934)."""
        return {"timestep": torch.arange(self.num_timesteps, dtype=torch.int64)}


def _cosine_logsnr(n, lo, hi):
    b = np.arctan(np.exp(-0.5 * hi))
    a = np.arctan(np.exp(-0.5 * lo)) - b
    t = torch.linspace(0, 1, n, dtype=torch.float32)
    return -2.0 * torch.log(torch.tan(a * t + b))


def _linear_logsnr(n, lo, hi):
    t = torch.linspace(0, 1, n, dtype=torch.float32)
    return hi + (lo - hi) * t


def _log1mexp(x):
    return torch.where(x > np.log(2), torch.log1p(-torch.exp(-x)), torch.log(-torch.expm1(-x)))


class ContinuousNoiseScheduler(NoiseScheduler):
    def __init__(self, num_scales: int, logsnr_schedule: str, loss_type: str = "l2", logsnr_min: float = -20,
                 logsnr_max: float = 20, **kwargs):
        super().__init__()
        if logsnr_schedule == "cosine":
            gammas = _cosine_logsnr(num_scales + 1, logsnr_min, logsnr_max)
        elif logsnr_schedule == "linear":
            gammas = _linear_logsnr(num_scales + 1, logsnr_min, logsnr_max)
        else:
            raise NotImplementedError(f"Noise schedule {logsnr_schedule} not implemented.")
        self.num_timesteps = num_scales
        self.loss_type = loss_type
        sigma2 = torch.sigmoid(-gammas)
        self.register_buffer("gammas", gammas.to(torch.float32))
        self.register_buffer("alphas", torch.sqrt(1.0 - sigma2).to(torch.float32))
        self.register_buffer("sigma2", sigma2.to(torch.float32))
        self.register_buffer("sqrt_sigma2", torch.sqrt(sigma2).to(torch.float32))

    def steps(self) -> int:
        return self.num_timesteps

    def continuous(self) -> bool:
        return True

    def logsnr(self, t: torch.Tensor) -> torch.Tensor:
        """fp32 index rule of the reference: long(t * num_timesteps), clamped (scheduler.py:518-522)."""
        idx = torch.clamp((t * self.num_timesteps).to(torch.long), 0, self.num_timesteps)
        return self.gammas.to(t.device).gather(-1, idx)

    def _lambda_tables(self, N: int):
        """Per loop index, on the host: the index arithmetic is fp32 and must round like the
        reference's (differs from integer arithmetic at a few indices)."""
        g = self.gammas.detach().cpu()
        i = torch.arange(N)                                          # int64, like torch.tensor([idx]*B)
        clamp = lambda u: torch.clamp((u * self.num_timesteps).to(torch.long), 0, self.num_timesteps)
        lam_s = g.gather(-1, clamp(i / N))                           # fp32 true-divide (ddpm.py:937-944)
        lam_t = g.gather(-1, clamp((i + 1) / N))
        return lam_s, lam_t, (i / N)

    def step_coefficients(self, prediction: str, num_sampling_steps: int, sampler: str = "ancestral"):
        lam_s, lam_t, _ = self._lambda_tables(num_sampling_steps)
        sig = torch.sigmoid
        c = torch.zeros(num_sampling_steps, 8, dtype=torch.float32)
        if prediction == "v":        # x0 = a*x - b*o (scheduler.py:536-544)
            c[:, 0], c[:, 1], form = torch.sqrt(sig(lam_t)), torch.sqrt(sig(-lam_t)), 0
        elif prediction == "epsilon":   # x0 = a*(x - o*b) (scheduler.py:524-534)
            c[:, 0], c[:, 1], form = torch.sqrt(1.0 + torch.exp(-lam_t)), torch.rsqrt(1.0 + torch.exp(lam_t)), 1
        else:
            raise NotImplementedError(prediction)
        if sampler == "ancestral":   # q_posterior mean + fixed-large variance (scheduler.py:414-494)
            alpha_s = torch.sqrt(sig(lam_s))
            alpha_st = torch.sqrt((1.0 + torch.exp(-lam_t)) / (1.0 + torch.exp(-lam_s)))
            r = torch.exp(lam_t - lam_s)
            one_minus_r = -torch.expm1(lam_t - lam_s)
            c[:, 2] = one_minus_r * alpha_s          # on x0
            c[:, 3] = r * alpha_st                   # on x_t
            logvar = _log1mexp(lam_s - lam_t) + torch.nn.functional.logsigmoid(-lam_t)
            c[:, 4] = torch.exp(0.5 * logvar)
        elif sampler == "ddim":      # samplers/ddim.py:88-123
            c[:, 2] = torch.sqrt(sig(lam_s))         # alpha_s on x0
            c[:, 3] = torch.sqrt(sig(-lam_s))        # stdv_s on eps
            c[:, 5] = torch.sqrt(1.0 + torch.exp(lam_t))
            c[:, 6] = torch.rsqrt(1.0 + torch.exp(-lam_t))
        else:
            raise NotImplementedError(sampler)
        return c, form

    def network_time_tables(self, num_sampling_steps: int):
        lam_s, lam_t, t = self._lambda_tables(num_sampling_steps)
        return {"timestep": t.to(torch.float32), "logsnr_t": lam_t, "logsnr_s": lam_s}


class DiscreteRectifiedFlowNoiseScheduler(NoiseScheduler):
    def __init__(self, steps: int, max_time: float, **kwargs):
        super().__init__()
        self._steps, self._max_time, self._epsilon = steps, max_time, 1e-3

    def continuous(self) -> bool:
        return False

    def steps(self) -> int:
        return self._steps

    def network_time_tables(self, num_sampling_steps: int):
        return {"timestep": torch.arange(self._steps, dtype=torch.int64)}

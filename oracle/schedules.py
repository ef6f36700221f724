"""Noise-schedule tables (oracle).  TEST INFRASTRUCTURE ONLY.

Restates xdiffusion/scheduler.py: DiscreteNoiseScheduler tables (:127-224),
the "fixed large" log-variance (:238-258), the continuous cosine logSNR table
(:21-25, :351-399) and its fp32 index rule (:518-522, diffusion/ddpm.py:936-954),
and the rectified-flow time grid (samplers/rectified_flow.py:46-58).
"""
import math

import numpy as np
import torch


def discrete_tables(num_scales=1000, schedule_type="linear", min_beta=1e-4, max_beta=2e-2):
    """fp64 construction, cast to fp32 at the end (scheduler.py:43-48,161-224)."""
    T = num_scales
    if schedule_type == "linear":
        scale = 1000 / T
        betas = torch.linspace(scale * min_beta, scale * max_beta, T, dtype=torch.float64)
    elif schedule_type == "cosine":
        x = torch.linspace(0, T, T + 1, dtype=torch.float64)
        ac = torch.cos(((x / T) + 0.008) / 1.008 * torch.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    elif schedule_type == "quadratic":
        scale = 1000 / T
        betas = torch.linspace((scale * min_beta) ** 0.5, (scale * max_beta) ** 0.5, T,
                               dtype=torch.float64) ** 2
    elif schedule_type == "sigmoid":
        scale = 1000 / T
        b = torch.linspace(-6, 6, T, dtype=torch.float64)
        betas = torch.sigmoid(b) * (scale * max_beta - scale * min_beta) + scale * min_beta
    else:
        raise NotImplementedError(schedule_type)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, 0)
    ac_prev = torch.cat([torch.ones(1, dtype=torch.float64), ac[:-1]])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    t64 = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
        "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
    }
    return {k: v.to(torch.float32) for k, v in t64.items()}


def fixed_large_logvar(tables):
    """log(cat(posterior_variance[1], betas[1:])) in fp32 on the fp32 buffers
    (scheduler.py:244-254)."""
    return torch.log(torch.cat([tables["posterior_variance"][1:2], tables["betas"][1:]]))


def cosine_logsnr_table(num_scales=1024, logsnr_min=-20.0, logsnr_max=20.0):
    """gammas of ContinuousNoiseScheduler: num_scales+1 entries (scheduler.py:21-25,366)."""
    b = np.arctan(np.exp(-0.5 * logsnr_max))
    a = np.arctan(np.exp(-0.5 * logsnr_min)) - b
    t = torch.linspace(0, 1, num_scales + 1, dtype=torch.float32)
    return -2.0 * torch.log(torch.tan(a * t + b))


def linear_logsnr_table(num_scales, logsnr_min, logsnr_max):
    t = torch.linspace(0, 1, num_scales + 1, dtype=torch.float32)
    return logsnr_max + (logsnr_min - logsnr_max) * t


def continuous_indices(i, num_sampling_steps, num_timesteps):
    """(idx_s, idx_t) for loop index i: fp32 true-divide of an int64 tensor, fp32
    multiply, truncation to long, clamp (ddpm.py:934-944, scheduler.py:518-522)."""
    t = torch.tensor([i])
    idx_s = torch.clamp(((t / num_sampling_steps) * num_timesteps).to(torch.long), 0, num_timesteps)
    idx_t = torch.clamp((((t + 1) / num_sampling_steps) * num_timesteps).to(torch.long), 0,
                        num_timesteps)
    return int(idx_s), int(idx_t)


def continuous_time(i, num_sampling_steps):
    """context["timestep"] for the continuous scheduler: fp32(i / N) (ddpm.py:954)."""
    return float((torch.tensor([i]) / num_sampling_steps)[0])


def rectified_flow_time(i, N=1000, T=1.0, eps=1e-3):
    """Network time for loop index i (rectified_flow.py:52-57): python double, then
    multiplied into an fp32 ones tensor."""
    k = N - (i + 1)
    num_t = k / N * (T - eps) + eps
    return float((torch.ones(1) * num_t)[0])

"""Per-timestep sampler updates (oracle).  TEST INFRASTRUCTURE ONLY.

Closed-form restatements of one reverse-process step, evaluated in fp32 in the
same operation order as the reference so they are bit-exact against it on CPU:
  * DDPM ancestral, discrete tables        samplers/ancestral.py:21-72,189-267
                                           scheduler.py:238-324
  * DDPM ancestral, continuous logSNR      scheduler.py:414-494,524-544
  * DDIM, continuous logSNR                samplers/ddim.py:43-123
  * rectified-flow Euler                   samplers/rectified_flow.py:46-84
  * dynamic thresholding                   utils.py:379-396
  * classifier-free guidance combine       samplers/ancestral.py:229-231
"""
import math

import torch
import torch.nn.functional as F


def cfg_combine(o_cond, o_uncond, w):
    return o_uncond + w * (o_cond - o_uncond)


def dynamic_threshold(x0, p, c):
    n = x0.shape[0]
    s = torch.quantile(x0.abs().reshape(n, -1), p, dim=-1)
    s = torch.clamp(s, min=1, max=c)
    y = torch.clip(x0.reshape(n, -1).T, -s, s) / s
    return y.T.reshape(x0.shape)


def _clip(x0, threshold):
    if threshold is None:
        return torch.clamp(x0, -1.0, 1.0)
    return dynamic_threshold(x0, threshold[0], threshold[1])


def ancestral_discrete(x, o, z, i, tables, logvar_table, prediction="epsilon", threshold=None):
    """x_{t-1} from x_t (x), network output (o) and noise (z) at loop index i."""
    if prediction == "epsilon":
        x0 = tables["sqrt_recip_alphas_cumprod"][i] * x - tables["sqrt_recipm1_alphas_cumprod"][i] * o
    elif prediction == "v":
        x0 = tables["sqrt_alphas_cumprod"][i] * x - tables["sqrt_one_minus_alphas_cumprod"][i] * o
    else:
        raise NotImplementedError(prediction)
    x0 = _clip(x0, threshold)
    mean = tables["posterior_mean_coef1"][i] * x0 + tables["posterior_mean_coef2"][i] * x
    if i == 0:
        return x0
    return mean + torch.exp(0.5 * logvar_table[i]) * z


def _log1mexp(x):
    return torch.where(x > math.log(2), torch.log1p(-torch.exp(-x)), torch.log(-torch.expm1(-x)))


def x0_from_v_continuous(x, o, lam_t):
    return torch.sqrt(torch.sigmoid(lam_t)) * x - torch.sqrt(torch.sigmoid(-lam_t)) * o


def x0_from_eps_continuous(x, o, lam_t):
    return torch.sqrt(1.0 + torch.exp(-lam_t)) * (x - o * torch.rsqrt(1.0 + torch.exp(lam_t)))


def ancestral_continuous(x, o, z, i, lam_s, lam_t, prediction="v", threshold=None):
    """lam_s / lam_t are fp32 0-d tensors = gammas[idx_s] / gammas[idx_t]."""
    lam_s = torch.as_tensor(lam_s, dtype=torch.float32)
    lam_t = torch.as_tensor(lam_t, dtype=torch.float32)
    if prediction == "v":
        x0 = x0_from_v_continuous(x, o, lam_t)
    else:
        x0 = x0_from_eps_continuous(x, o, lam_t)
    x0 = _clip(x0, threshold)
    alpha_s = torch.sqrt(torch.sigmoid(lam_s))
    alpha_st = torch.sqrt((1.0 + torch.exp(-lam_t)) / (1.0 + torch.exp(-lam_s)))
    r = torch.exp(lam_t - lam_s)
    one_minus_r = -torch.expm1(lam_t - lam_s)
    mean = r * alpha_st * x + one_minus_r * alpha_s * x0
    logvar = _log1mexp(lam_s - lam_t) + F.logsigmoid(-lam_t)
    if i == 0:
        return x0
    return mean + torch.exp(0.5 * logvar) * z


def ddim_continuous(x, o, i, lam_s, lam_t, prediction="v", threshold=None):
    lam_s = torch.as_tensor(lam_s, dtype=torch.float32)
    lam_t = torch.as_tensor(lam_t, dtype=torch.float32)
    if prediction == "v":
        x0u = x0_from_v_continuous(x, o, lam_t)
        # epsilon from the UNclipped x0 (ddim.py:88-93, scheduler.py:553-558)
        e = torch.sqrt(1.0 + torch.exp(lam_t)) * (x - x0u * torch.rsqrt(1.0 + torch.exp(-lam_t)))
    else:
        x0u = x0_from_eps_continuous(x, o, lam_t)
        e = o
    x0 = _clip(x0u, threshold)
    if i == 0:
        return x0
    return torch.sqrt(torch.sigmoid(lam_s)) * x0 + torch.sqrt(torch.sigmoid(-lam_s)) * e


def euler_flow(x, o, N=1000):
    """sigma_t == 0 => x + v * dt; the reference still evaluates
    pred + 0.0/(..)*(..) and adds 0.0*sqrt(dt)*randn (rectified_flow.py:76-84)."""
    return x + o * (1.0 / N)


def unnormalize(x):
    """utils.py:62-64"""
    return (torch.clamp(x, -1.0, 1.0) + 1) * 0.5


# ----------------------------------------------------------------------------- EDM (samplers/edm.py:10-137)
def edm_time_steps(num_steps, sigma_min, sigma_max, rho):
    """samplers/edm.py:49-62 (round_sigma of EDMPrecond is the identity, score_networks/edm.py:695-696)."""
    idx = torch.arange(num_steps, dtype=torch.float64)
    t = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([torch.as_tensor(t), torch.zeros_like(t[:1])])


def edm_precond(raw_model, x, sigma, sigma_data=0.5):
    """EDMPrecond.forward (score_networks/edm.py:663-693), fp32."""
    x = x.to(torch.float32)
    sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / (sigma ** 2 + sigma_data ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + sigma ** 2).sqrt()
    c_noise = sigma.log() / 4
    f = raw_model((c_in * x).to(torch.float32), c_noise.flatten())
    return c_skip * x + c_out * f.to(torch.float32)


def edm_sample(raw_model, latents, num_steps=18, sigma_min=0.002, sigma_max=80.0, rho=7, S_churn=0, S_min=0,
               S_max=float("inf"), S_noise=1, sigma_data=0.5, noise=None, trace=None):
    """StochasticSampler.p_sample_loop + p_sample (samplers/edm.py:36-137); ``noise[i]`` replaces randn_like at step i."""
    import numpy as np
    t_steps = edm_time_steps(num_steps, sigma_min, sigma_max, rho)
    x_next = latents.to(torch.float64) * t_steps[0]
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        gamma = min(S_churn / num_steps, np.sqrt(2) - 1) if S_min <= t_cur <= S_max else 0
        t_hat = torch.as_tensor(t_cur + gamma * t_cur)
        z = noise[i] if noise is not None else torch.randn_like(x_cur)
        x_hat = x_cur + (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise * z
        denoised = edm_precond(raw_model, x_hat, t_hat, sigma_data).to(torch.float64)
        d_cur = (x_hat - denoised) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            denoised = edm_precond(raw_model, x_next, t_next, sigma_data).to(torch.float64)
            d_prime = (x_next - denoised) / t_next
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
        if trace is not None:
            trace.append(x_next)
    return x_next

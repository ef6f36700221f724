"""EDM stochastic sampler (Karras et al. 2022, Algorithm 2) on the fused fp64 step kernels.

Drop-in for ``xdiffusion.samplers.edm.StochasticSampler`` (reference: samplers/edm.py:10-137): same constructor, same
``p_sample_loop(diffusion_model, latents, class_labels)``.  The state is fp64 like the reference's.  Per step the
reference launches ~25 elementwise kernels around two network evaluations and moves the fp64 state ~20 times through
memory; here one kernel per network evaluation folds the EDMPrecond output arithmetic (c_skip x + c_out F), the Euler /
Heun update and the scaling of the NEXT network input (csrc/step.cu: xd_edm_step, xd_edm_prepare).  Every floating-point
operation is the reference's, un-fused and in its order: given the same raw network outputs the state is bit-identical
(tests/test_edm_*.py, fixture from the real reference).  The time-step discretisation is evaluated with the same torch
fp64 expressions on the host (18-35 scalars).
"""
from typing import Optional

import numpy as np
import torch


class StochasticSampler:
    def __init__(self, num_steps: int = 18, sigma_min: float = 0.002, sigma_max: float = 80, rho: float = 7,
                 S_churn: float = 0, S_min: float = 0, S_max: float = float("inf"), S_noise: float = 1):
        self._num_steps, self._sigma_min, self._sigma_max, self._rho = num_steps, sigma_min, sigma_max, rho
        self._S_churn, self._S_min, self._S_max, self._S_noise = S_churn, S_min, S_max, S_noise

    def time_steps(self, score_network) -> torch.Tensor:
        """t_0 .. t_N (fp64, t_N = 0), reference samplers/edm.py:44-62."""
        sigma_min = max(self._sigma_min, score_network.sigma_min)
        sigma_max = min(self._sigma_max, score_network.sigma_max)
        idx = torch.arange(self._num_steps, dtype=torch.float64)
        t = (sigma_max ** (1 / self._rho)
             + idx / (self._num_steps - 1) * (sigma_min ** (1 / self._rho) - sigma_max ** (1 / self._rho))) ** self._rho
        return torch.cat([score_network.round_sigma(t), torch.zeros_like(t[:1])])

    @torch.no_grad()
    def p_sample_loop(self, diffusion_model, latents: torch.Tensor, class_labels: Optional[torch.Tensor] = None,
                      noise=None, trace=None):
        """latents fp32/fp64 (B, C, H, W) on the GPU -> fp64 samples.  ``noise``: optional list of fp64 tensors, one per
        step, replacing ``randn_like`` of the churn branch (parity runs); ``trace`` (list) receives the state after every
        step."""
        net = diffusion_model._score_network
        if latents.device.type != "cuda":
            raise RuntimeError("xdiffusion_b200 runs on CUDA (sm_100a) only")
        t_steps = self.time_steps(net)
        x = (latents.to(torch.float64) * t_steps[0].to(latents.device)).contiguous()
        n = self._num_steps
        x_hat, x_mid, d_cur = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        xin = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        step = torch.ops.xdb200.edm_step
        have_xin = False                                  # xin already holds c_in(t_hat) * float(x) from the previous kernel
        for i in range(n):
            t_cur, t_next = t_steps[i], t_steps[i + 1]
            gamma = min(self._S_churn / n, np.sqrt(2) - 1) if self._S_min <= t_cur <= self._S_max else 0
            t_hat = net.round_sigma(t_cur + gamma * t_cur)
            c_skip, c_out, c_in, c_noise = net.precond_scalars(t_hat)
            if gamma > 0:                                 # temporary noise increase (edm.py:108-120)
                z = noise[i].to(x.device, torch.float64) if noise is not None else torch.randn_like(x)
                c = float((t_hat ** 2 - t_cur ** 2).sqrt() * self._S_noise)
                torch.ops.xdb200.edm_prepare(x, z.contiguous(), c, x_hat, c_in, xin)
                cur = x_hat
            else:
                cur = x
                if not have_xin:
                    torch.ops.xdb200.edm_prepare(x, None, 0.0, None, c_in, xin)
            f1 = net.raw(xin, c_noise, class_labels)
            h = float(t_next - t_hat)
            heun = i < n - 1
            c_skip2, c_out2, c_in2, c_noise2 = net.precond_scalars(t_next) if heun else (0.0, 0.0, 0.0, 0.0)
            # Euler: d_cur, x_mid = cur + h d_cur, and the input of the second evaluation c_in(t_next) * float(x_mid)
            step(0, cur, None, None, f1, d_cur, x_mid, None, xin if heun else None, float(t_hat), h, c_skip, c_out, c_in2)
            if heun:                                      # 2nd-order correction (edm.py:131-136)
                f2 = net.raw(xin, c_noise2, class_labels)
                # ... and the input of the NEXT step's first evaluation (its t_hat = t_next when there is no churn)
                nxt_gamma = min(self._S_churn / n, np.sqrt(2) - 1) if self._S_min <= t_next <= self._S_max else 0
                step(1, cur, x_mid, d_cur, f2, None, x, None, xin if nxt_gamma == 0 else None, float(t_next), h,
                     c_skip2, c_out2, c_in2)
                have_xin = nxt_gamma == 0
            else:
                x, x_mid = x_mid, x
                have_xin = False
            if trace is not None:
                trace.append(x.clone())
        return x

"""EDM preconditioning wrapper (reference: score_networks/edm.py:635-697).

``EDMPrecond`` keeps the reference constructor and ``forward(x, sigma, class_labels)``; the sampler does not call
``forward`` but ``raw`` + ``precond_scalars`` so that the output arithmetic runs inside the fused step kernel.  The four
scalars are evaluated with the reference's own fp32 tensor expressions (one value per step; on the host).

The raw networks of the reference's EDM configs (``SongUNet`` = DDPM++ / NCSN++, ``DhariwalUNet`` = ADM) are NOT built on
the B200 kernels (they need a single-head attention with head dim = C and resampling convolutions the benchmark networks
do not use); instantiating them raises NotImplementedError like every other out-of-scope network.  Any module with the raw
signature ``model(x, noise_labels, class_labels=None)`` can be wrapped.
"""
import torch

from ..utils import instantiate_from_config


class EDMPrecond(torch.nn.Module):
    def __init__(self, img_resolution, img_channels, label_dim=0, use_fp16=False, sigma_min=0, sigma_max=float("inf"),
                 sigma_data=0.5, model_type="DhariwalUNet", **model_kwargs):
        super().__init__()
        if use_fp16:
            raise NotImplementedError("use_fp16")
        self.img_resolution, self.img_channels, self.label_dim = img_resolution, img_channels, label_dim
        self.sigma_min, self.sigma_max, self.sigma_data = sigma_min, sigma_max, sigma_data
        self.model = instantiate_from_config(model_kwargs["model"])

    def precond_scalars(self, sigma):
        """(c_skip, c_out, c_in, c_noise) as Python floats holding the reference's fp32 values (edm.py:682-685)."""
        s = torch.as_tensor(sigma).to(torch.float32).reshape(-1, 1, 1, 1).cpu()
        sd = self.sigma_data
        c_skip = sd ** 2 / (s ** 2 + sd ** 2)
        c_out = s * sd / (s ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + s ** 2).sqrt()
        c_noise = s.log() / 4
        return float(c_skip), float(c_out), float(c_in), float(c_noise)

    def _labels(self, x, class_labels):
        if self.label_dim == 0:
            return None
        if class_labels is None:
            return torch.zeros([1, self.label_dim], device=x.device)
        return class_labels.to(torch.float32).reshape(-1, self.label_dim)

    def raw(self, xin, c_noise: float, class_labels=None):
        """F = model(c_in * x, c_noise): fp32 in, fp32 out."""
        noise_labels = torch.full((1,), c_noise, device=xin.device, dtype=torch.float32)
        out = self.model(xin, noise_labels, class_labels=self._labels(xin, class_labels))
        assert out.dtype == torch.float32
        return out.contiguous()

    def forward(self, x, sigma, class_labels=None, force_fp32=False, **model_kwargs):
        """D(x; sigma) = c_skip x + c_out F(c_in x; c_noise)  (one sigma for the batch, as the samplers call it)."""
        c_skip, c_out, c_in, c_noise = self.precond_scalars(sigma)
        x64 = x.to(torch.float64).contiguous()
        xin = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        torch.ops.xdb200.edm_prepare(x64, None, 0.0, None, c_in, xin)
        f = self.raw(xin, c_noise, class_labels)
        den = torch.empty_like(xin)
        torch.ops.xdb200.edm_step(2, x64, None, None, f, None, None, den, None, 1.0, 0.0, c_skip, c_out, 0.0)
        return den

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)

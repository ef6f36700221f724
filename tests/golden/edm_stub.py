"""A stand-in for the raw EDM score network with exactly reproducible arithmetic (one fp32 multiply, one fp32 multiply-free
add per element: identical on CPU and GPU), used to pin the EDM SAMPLER and PRECONDITIONING arithmetic against the real
reference without the DDPM++ network (which is not built).  Signature of the reference's raw models
(score_networks/edm.py: SongUNet.forward(x, noise_labels, class_labels, augment_labels=None))."""
import torch


class AffineStub(torch.nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.scale = torch.nn.Parameter(torch.tensor(0.5), requires_grad=False)

    def forward(self, x, noise_labels, class_labels=None, augment_labels=None):
        return x * self.scale - 0.25 * noise_labels.reshape(-1, 1, 1, 1).to(x.dtype)

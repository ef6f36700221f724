"""xdiffusion_b200 -- the sampling hot path of xdiffusion on hand-written sm_100a kernels.

Drop-in for ``GaussianDiffusion_DDPM(config).sample()`` and the score-network / sampler /
scheduler classes it instantiates from YAML ``target:`` paths (``xdiffusion.`` prefixes are
resolved to this package).  No CPU or PyTorch-eager fallback exists.
"""
__version__ = "0.1.0"

"""EDM network fixture (the raw network): the REAL reference's EDMPrecond + SongUNet (DDPM++) of
``configs/image/mnist/edm.yaml`` with seeded synthetic weights (the reference initialises the second conv of every block, the
attention projections and the output conv to ~1e-5, which would make parity vacuous).

    python tests/golden/make_edm_net.py            # writes tests/golden/edm_net.pt

Stored: the preconditioned denoiser D(x; sigma) at a few noise levels, the raw network output F at one, and a short run of the
reference's StochasticSampler (4 steps = 7 network evaluations).  Only runs in the authoring container.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_bootstrap  # noqa: E402
from oracle.weights import synth_state_dict  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref_bootstrap.bootstrap()
    from xdiffusion.samplers.edm import StochasticSampler
    from xdiffusion.score_networks.edm import EDMPrecond
    from xdiffusion.utils import load_yaml
    cfg = load_yaml(os.path.join(ref_bootstrap.REFERENCE_ROOT, "configs/image/mnist/edm.yaml")).to_dict()
    torch.manual_seed(0)
    net = EDMPrecond(**cfg["diffusion"]["score_network"]["params"]).eval()
    manifest = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth_state_dict(manifest, seed=0)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all("resample_filter" in k for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(321)
    B = 2
    out = {"manifest": manifest, "config": cfg, "denoise": []}
    with torch.no_grad():
        for sigma in (80.0, 3.7, 0.3, 0.002):
            x = torch.randn(B, 1, 32, 32, generator=g) * sigma
            d = net(x, torch.tensor(sigma))
            out["denoise"].append({"sigma": sigma, "x": x, "D": d})
        x = torch.randn(B, 1, 32, 32, generator=g)
        out["raw"] = {"x": x, "c_noise": torch.tensor([0.41]), "F": net.model(x, torch.tensor([0.41]), class_labels=None)}
        latents = torch.randn(B, 1, 32, 32, generator=g)
        sampler = StochasticSampler(num_steps=4)
        states = []
        orig = sampler.p_sample

        def rec(*a, **k):
            r = orig(*a, **k)
            states.append(r.clone())
            return r

        sampler.p_sample = rec
        final = sampler.p_sample_loop(diffusion_model=types.SimpleNamespace(_score_network=net), latents=latents.clone())
    out["sampler"] = {"num_steps": 4, "latents": latents, "states": states}
    torch.save(out, os.path.join(HERE, "edm_net.pt"))
    print("edm_net written", [float(r["D"].abs().mean()) for r in out["denoise"]], float(final.abs().mean()))


if __name__ == "__main__":
    main()

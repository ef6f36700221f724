"""EDM sampler fixture: the REAL reference's StochasticSampler + EDMPrecond (samplers/edm.py,
score_networks/edm.py) around a stub raw network with exactly reproducible arithmetic (tests/golden/edm_stub.py).

    python tests/golden/make_edm.py            # writes tests/golden/edm.pt

Two runs of 6 steps at batch 2 x 1 x 32 x 32: deterministic (S_churn = 0) and with churn (S_churn = 4, injected fp64 noise).
The fixture stores the latents, the noise and the fp64 state after every step.  Only runs in the authoring container.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_bootstrap  # noqa: E402
from tests.golden.make_golden import NoiseFeeder  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref_bootstrap.bootstrap()
    from xdiffusion.samplers.edm import StochasticSampler
    from xdiffusion.score_networks.edm import EDMPrecond
    net = EDMPrecond(img_resolution=32, img_channels=1, label_dim=0, sigma_data=0.5,
                     model={"target": "tests.golden.edm_stub.AffineStub", "params": {}}).eval()
    dm = types.SimpleNamespace(_score_network=net)
    g = torch.Generator().manual_seed(77)
    latents = torch.randn(2, 1, 32, 32, generator=g)
    out = {"latents": latents, "runs": {}}
    for name, churn in (("plain", 0.0), ("churn", 4.0)):
        n = 6
        sampler = StochasticSampler(num_steps=n, S_churn=churn)
        zs = [torch.randn(latents.shape, generator=g, dtype=torch.float64) for _ in range(n)]
        states = []
        orig = sampler.p_sample

        def rec(*a, _orig=orig, **k):
            r = _orig(*a, **k)
            states.append(r.clone())
            return r

        sampler.p_sample = rec
        with NoiseFeeder() as feeder:
            feeder.queue = [z.clone() for z in zs]          # one randn_like per step, in order
            final = sampler.p_sample_loop(diffusion_model=dm, latents=latents.clone(), class_labels=None)
        assert final.dtype == torch.float64 and torch.equal(final, states[-1])
        out["runs"][name] = {"num_steps": n, "S_churn": churn, "noise": zs, "states": states}
        print(name, "final |x| mean", float(final.abs().mean()))
    torch.save(out, os.path.join(HERE, "edm.pt"))


if __name__ == "__main__":
    main()

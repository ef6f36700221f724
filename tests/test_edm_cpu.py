"""SURVEY 8(f2), host side: the oracle's restatement of the EDM stochastic sampler + EDMPrecond reproduces, bit for bit, the
fp64 states the REAL reference produced (tests/golden/make_edm.py; stub raw network with exactly reproducible arithmetic),
and the product's host-side pieces (time-step discretisation, preconditioning scalars) equal the oracle's."""
import os

import torch

from oracle import samplers as osamplers
from tests.conftest import GOLDEN
from tests.golden.edm_stub import AffineStub


def _fx():
    return torch.load(os.path.join(GOLDEN, "edm.pt"), weights_only=False)


def test_oracle_edm_sampler_bit_exact_vs_reference():
    fx = _fx()
    stub = AffineStub()
    for name, run in fx["runs"].items():
        trace = []
        with torch.no_grad():
            osamplers.edm_sample(lambda x, c: stub(x, c), fx["latents"], num_steps=run["num_steps"], S_churn=run["S_churn"],
                                 noise=run["noise"], trace=trace)
        assert len(trace) == len(run["states"])
        for i, (a, b) in enumerate(zip(trace, run["states"])):
            assert a.dtype == torch.float64 and torch.equal(a, b), (name, i, float((a - b).abs().max()))


def test_product_time_steps_and_scalars_match_oracle():
    from xdiffusion_b200.samplers.edm import StochasticSampler
    from xdiffusion_b200.score_networks.edm import EDMPrecond
    net = EDMPrecond(img_resolution=32, img_channels=1, model={"target": "tests.golden.edm_stub.AffineStub", "params": {}})
    for n in (6, 18, 35):
        t = StochasticSampler(num_steps=n).time_steps(net)
        assert torch.equal(t, osamplers.edm_time_steps(n, 0.002, 80, 7)) and float(t[-1]) == 0.0 and t.dtype == torch.float64
    t = StochasticSampler(num_steps=6).time_steps(net)
    for sigma in t[:-1]:
        s = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
        want = (0.25 / (s ** 2 + 0.25), s * 0.5 / (s ** 2 + 0.25).sqrt(), 1 / (0.25 + s ** 2).sqrt(), s.log() / 4)
        assert net.precond_scalars(sigma) == tuple(float(w) for w in want)


def _net_fx():
    return torch.load(os.path.join(GOLDEN, "edm_net.pt"), weights_only=False)


def test_oracle_songunet_and_precond_match_reference():
    """The oracle's DDPM++ network (oracle.nets.songunet_forward) + EDMPrecond against the real reference
    (tests/golden/make_edm_net.py): raw output, denoiser at four noise levels, and a 4-step Heun run."""
    from oracle import nets
    from oracle.weights import synth_state_dict
    fx = _net_fx()
    sd = synth_state_dict(fx["manifest"], seed=0)
    p = fx["config"]["diffusion"]["score_network"]["params"]["model"]["params"]
    raw = lambda x, c: nets.songunet_forward(sd, p, x, c)
    with torch.no_grad():
        r = fx["raw"]
        f = raw(r["x"], r["c_noise"])
        assert float((f - r["F"]).norm() / r["F"].norm()) < 1e-5
        for rec in fx["denoise"]:
            d = osamplers.edm_precond(raw, rec["x"], torch.tensor(rec["sigma"]))
            assert float((d - rec["D"]).norm() / rec["D"].norm()) < 1e-5, rec["sigma"]
        sm = fx["sampler"]
        trace = []
        osamplers.edm_sample(raw, sm["latents"], num_steps=sm["num_steps"], noise=[torch.zeros_like(sm["latents"], dtype=torch.float64)] * 4,
                             trace=trace)
        for a, b in zip(trace, sm["states"]):
            assert float((a - b).norm() / b.norm()) < 1e-5


def test_product_songunet_state_dict_matches_reference_manifest():
    """The product's DDPM++ network built from the reference's own configs/image/mnist/edm.yaml has exactly the reference's
    state_dict keys and shapes (incl. the ``resample_filter`` buffers), and the reference's values in those buffers."""
    from xdiffusion_b200.diffusion.edm import GaussianDiffusion_EDM
    from xdiffusion_b200.utils import DotConfig
    fx = _net_fx()
    m = GaussianDiffusion_EDM(DotConfig(fx["config"]))
    sd = m._score_network.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == fx["manifest"]
    for k, v in sd.items():
        if k.endswith("resample_filter"):
            assert torch.equal(v, torch.full((1, 1, 2, 2), 0.25))

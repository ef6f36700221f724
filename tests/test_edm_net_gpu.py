"""SURVEY 8(f2), the raw network: the DDPM++ (SongUNet) network + EDMPrecond on the B200 kernels against the outputs of the
REAL reference (tests/golden/make_edm_net.py, configs/image/mnist/edm.yaml with seeded synthetic weights)."""
import os

import pytest
import torch

from oracle.weights import synth_state_dict
from tests.conftest import GOLDEN
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_TOL = 2e-2           # the repo-wide per-evaluation tolerance of the bf16 tensor-core path


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN, "edm_net.pt"), weights_only=False)


@pytest.fixture(scope="module")
def model(fx):
    from xdiffusion_b200.diffusion.edm import GaussianDiffusion_EDM
    from xdiffusion_b200.utils import DotConfig
    m = GaussianDiffusion_EDM(DotConfig(fx["config"]))
    missing, unexpected = m._score_network.load_state_dict(synth_state_dict(fx["manifest"], seed=0), strict=False)
    assert not unexpected and all("resample_filter" in k for k in missing), (missing, unexpected)
    return m.to(DEV).eval()


def test_raw_network_matches_reference(fx, model):
    r = fx["raw"]
    with torch.no_grad():
        f = model._score_network.model(r["x"].to(DEV), r["c_noise"].to(DEV), class_labels=None)
    assert f.dtype == torch.float32 and rel_l2(f, r["F"]) < BF16_TOL, rel_l2(f, r["F"])


def test_denoiser_matches_reference_at_every_noise_level(fx, model):
    """D(x; sigma) = c_skip x + c_out F(c_in x; c_noise); the error of F is also bounded on its own (at small sigma D ~ x
    hides it)."""
    net = model._score_network
    for rec in fx["denoise"]:
        sigma = torch.tensor(rec["sigma"])
        with torch.no_grad():
            d = net(rec["x"].to(DEV), sigma)
        assert rel_l2(d, rec["D"]) < BF16_TOL, (rec["sigma"], rel_l2(d, rec["D"]))
        c_skip, c_out, _, _ = net.precond_scalars(sigma)
        f, f_ref = (d.cpu() - c_skip * rec["x"]) / c_out, (rec["D"] - c_skip * rec["x"]) / c_out
        if c_skip < 0.9:                                          # (cancellation-free levels only)
            assert rel_l2(f, f_ref) < BF16_TOL, (rec["sigma"], rel_l2(f, f_ref))


def test_heun_sampler_states_match_reference(fx, model):
    from xdiffusion_b200.samplers.edm import StochasticSampler
    sm = fx["sampler"]
    trace = []
    x = StochasticSampler(num_steps=sm["num_steps"]).p_sample_loop(model, sm["latents"].to(DEV), trace=trace)
    assert x.dtype == torch.float64 and len(trace) == len(sm["states"])
    for i, (a, b) in enumerate(zip(trace, sm["states"])):
        assert rel_l2(a, b) < BF16_TOL, (i, rel_l2(a, b))
    out, _ = model.sample(num_samples=2, initial_noise=sm["latents"].to(DEV))      # configured sampler: 18 steps, runs
    assert out.shape == (2, 1, 32, 32) and bool(torch.isfinite(out).all()) and 0 <= float(out.min()) and float(out.max()) <= 1


def test_batch_rows_are_independent(fx, model):
    """A 64-image batch (the benchmark batch of the other UNet workloads) reproduces the fixture rows placed at its ends."""
    r = fx["raw"]
    g = torch.Generator().manual_seed(9)
    x = torch.randn(64, 1, 32, 32, generator=g)
    x[0], x[-1] = r["x"][0], r["x"][1]
    with torch.no_grad():
        f = model._score_network.model(x.to(DEV), r["c_noise"].to(DEV), class_labels=None).cpu()
    assert rel_l2(f[[0, -1]], r["F"]) < BF16_TOL, rel_l2(f[[0, -1]], r["F"])

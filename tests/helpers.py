"""Shared test plumbing: build the product model from a golden fixture and load the seeded weights."""
import torch

from oracle import nets as onets
from oracle.loop import OracleModel
from oracle.weights import synth_state_dict


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def fixture_state_dict(fx):
    sd = synth_state_dict(fx["manifest"], seed=0)
    if "pos_embed" in fx["manifest"]:
        sd["pos_embed"] = onets.sincos_pos_embed_2d(384, 4, 16 if fx["kind"] == "dit" else 4)[None]
    return sd


def oracle_model(fx):
    return OracleModel(fx["kind"], fx["config"], fixture_state_dict(fx))


def product_model(fx, device="cuda"):
    from xdiffusion_b200.diffusion import GaussianDiffusion_DDPM
    from xdiffusion_b200.utils import DotConfig
    m = GaussianDiffusion_DDPM(DotConfig(fx["config"]))
    sd = {"_score_network." + k: v for k, v in fixture_state_dict(fx).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    learnable_missing = [k for k in missing if k.startswith("_score_network.") and "_context_transformers" not in k
                         and "pos_embed" not in k]
    assert not learnable_missing, learnable_missing
    return m.to(device).eval()

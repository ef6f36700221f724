"""LoRA / checkpoint on-ramp, host side: the product's module traversal lists the same layers in the same
order as the reference's ``_find_modules`` (fixture shapes come from the reference's own injection), and merging the
low-rank deltas into the base weights reproduces the reference's LoraInjected* forward -- checked with the CPU oracle on
the merged state dict against scores the REAL reference produced (tests/golden/make_lora.py)."""
import os

import pytest
import torch

from oracle import nets
from tests.conftest import GOLDEN
from tests.helpers import fixture_state_dict, rel_l2
from xdiffusion_b200 import lora
from xdiffusion_b200.utils import DotConfig


def _product_cpu(fx):
    from xdiffusion_b200.diffusion import GaussianDiffusion_DDPM
    m = GaussianDiffusion_DDPM(DotConfig(fx["config"]))
    m.load_state_dict({"_score_network." + k: v for k, v in fixture_state_dict(fx).items()}, strict=False)
    return m


@pytest.mark.parametrize("name", ["c1", "c7"])
def test_lora_merge_matches_reference_injection(name, golden):
    lf = torch.load(os.path.join(GOLDEN, f"lora_{name}.pt"), weights_only=False)
    fx = golden(name)
    m = _product_cpu(fx)
    assert lora.lora_shapes(m, r=lf["rank"]) == [(tuple(u), tuple(d)) for u, d in lf["shapes"]]
    p = fx["config"]["diffusion"]["score_network"]["params"]
    t, text = lf["ctx"]["timestep"], lf["ctx"].get("text_embeddings")

    def oracle_score():
        sd = {k[len("_score_network."):]: v for k, v in m.state_dict().items() if k.startswith("_score_network.")}
        with torch.no_grad():
            return nets.unet_forward(sd, p, lf["x"], t, text=text)

    assert rel_l2(oracle_score(), lf["score_base"]) < 1e-5
    names = lora.merge_lora_weights(m, lora.synth_lora(lf["shapes"], seed=lf["seed"]))
    assert len(names) == len(lf["shapes"])
    err = rel_l2(oracle_score(), lf["score_lora"])
    assert err < 1e-4, err                                          # merged weights == injected up(down(x)) branch
    assert rel_l2(lf["score_lora"], lf["score_base"]) > 0.1          # and the LoRA is not a no-op
    assert lora.remove_lora_weights(m) == len(names)
    assert rel_l2(oracle_score(), lf["score_base"]) < 1e-5


def test_lora_file_round_trip_and_errors(golden, tmp_path):
    fx = golden("c1")
    m = _product_cpu(fx)
    shapes = lora.lora_shapes(m, r=4)
    path = tmp_path / "lora.pt"
    torch.save(lora.synth_lora(shapes, seed=1), path)               # a list of Parameters, like save_lora_weights writes
    w = m._score_network.downs[0][0].in_layers[2].weight
    before, version = w.detach().clone(), w._version
    names = lora.load_lora_weights(m, str(path))
    assert names[0].endswith("downs.0.0.in_layers.2") and w._version > version and not torch.equal(w, before)
    with pytest.raises(ValueError):
        lora.merge_lora_weights(m, lora.synth_lora(shapes[:-1], seed=1))
    bad = lora.synth_lora(shapes, seed=1)
    bad[0] = torch.nn.Parameter(torch.zeros(7, 4, 1, 1))
    with pytest.raises(ValueError):
        lora.merge_lora_weights(m, bad)


def test_checkpoint_on_disk_loads_like_the_reference(golden, tmp_path):
    """diffusion/ddpm.py:795-814: torch.load(path)["model_state_dict"], strict=False, keys under `_score_network.`."""
    fx = golden("c1")
    src = _product_cpu(fx)
    path = tmp_path / "ckpt.pt"
    torch.save({"model_state_dict": src.state_dict(), "step": 123}, path)
    from xdiffusion_b200.diffusion import GaussianDiffusion_DDPM
    torch.manual_seed(7)
    dst = GaussianDiffusion_DDPM(DotConfig(fx["config"]))
    fp = dst._weights_fingerprint()
    dst.load_checkpoint(str(path))
    assert dst._weights_fingerprint() != fp                         # captured loops keyed on it are invalidated
    a, b = src.state_dict(), dst.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)

"""Host-side logic that needs no GPU: config plumbing, state_dict key compatibility with the
reference (fixture manifests come from the reference's own state_dict()), schedule tables and
per-step coefficient rows against the oracle."""
import pytest
import torch

from oracle import schedules as osched
from tests.helpers import fixture_state_dict
from xdiffusion_b200.utils import DotConfig, get_obj_from_str, instantiate_from_config, resolve_target


def _build(fx):
    from xdiffusion_b200.diffusion import GaussianDiffusion_DDPM
    return GaussianDiffusion_DDPM(DotConfig(fx["config"]))


def test_target_paths_resolve_to_dropins():
    assert resolve_target("xdiffusion.score_networks.dit.DiT") == "xdiffusion_b200.score_networks.dit.DiT"
    assert resolve_target("torch.nn.Identity") == "torch.nn.Identity"
    cls = get_obj_from_str("xdiffusion.samplers.ancestral.AncestralSampler")
    assert cls.__module__ == "xdiffusion_b200.samplers.ancestral"
    with pytest.raises(NotImplementedError):
        get_obj_from_str("xdiffusion.score_networks.flux.Flux")
    s = instantiate_from_config({"target": "xdiffusion.scheduler.DiscreteNoiseScheduler",
                                 "params": {"num_scales": 1000, "schedule_type": "linear", "loss_type": "l2"}})
    assert s.steps() == 1000 and not s.continuous()


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5", "c6", "c7", "c8"])
def test_state_dict_keys_match_reference(name, golden):
    fx = golden(name)
    try:
        m = _build(fx)
    except NotImplementedError as e:
        pytest.skip(f"not built yet: {e}")
    ours = {k[len("_score_network."):]: tuple(v.shape) for k, v in m.state_dict().items()
            if k.startswith("_score_network.")}
    ours = {k: v for k, v in ours.items() if not (k.startswith("_context_transformers.") and "._projections." in k)}
    assert ours == fx["manifest"]
    missing, unexpected = m.load_state_dict({"_score_network." + k: v for k, v in fixture_state_dict(fx).items()},
                                            strict=False)
    assert not unexpected


def test_scheduler_buffers_and_coefficients_match_oracle(golden):
    from xdiffusion_b200.scheduler import ContinuousNoiseScheduler, DiscreteNoiseScheduler
    kat = golden("kat")
    d = DiscreteNoiseScheduler("linear", 1000, "l2")
    for k, v in kat["discrete"].items():
        assert torch.equal(getattr(d, k), v), k
    c, form = d.step_coefficients("epsilon", 1000)
    assert form == 0 and torch.equal(c[:, 4], torch.exp(0.5 * kat["fixed_large_logvar"]))
    s = ContinuousNoiseScheduler(1024, "cosine", "l2", -20, 20)
    assert torch.equal(s.gammas, kat["gammas"])
    for N in (1024, 1000, 50):
        tabs = s.network_time_tables(N)
        assert torch.equal(tabs["logsnr_s"], kat[f"logsnr_s_{N}"])
        assert torch.equal(tabs["logsnr_t"], kat[f"logsnr_t_{N}"])
        assert torch.equal(tabs["timestep"], torch.arange(N) / N)
    # the fp32 index rule differs from integer arithmetic at a few indices
    s1000 = ContinuousNoiseScheduler(1000, "cosine", "l2", -20, 20)
    ls = s1000.network_time_tables(1000)["logsnr_s"]
    assert int((ls != s1000.gammas[torch.arange(1000)]).sum()) > 0


def test_pos_embed_matches_reference_table(golden):
    from xdiffusion_b200.layers.utils import get_2d_sincos_pos_embed
    kat = golden("kat")
    assert torch.equal(torch.from_numpy(get_2d_sincos_pos_embed(384, 4)).float(), kat["pos_dit"])
    assert torch.equal(torch.from_numpy(get_2d_sincos_pos_embed(384, 4, lewei_scale=(1.0,), base_size=4)).float(),
                       kat["pos_pixart"])


def test_flow_time_table_matches_oracle(golden):
    fx = golden("c3")
    m = _build(fx)
    t = m._time_tables(1000, "cpu")["timestep"]
    ref = torch.tensor([osched.rectified_flow_time(i) for i in range(1000)])
    assert torch.equal(t, ref)


def test_cpu_model_refuses_to_sample(golden):
    m = _build(golden("c2"))
    with pytest.raises(RuntimeError):
        m.sample(context={"classes": torch.tensor([1, 2])}, num_samples=2, num_sampling_steps=2)


def test_cfg_merge_lines_up_conditional_and_unconditional_rows():
    """Host logic of the one-forward classifier-free guidance (diffusion/ddpm.py: _DeviceLoop._merge): per-sample tensors are
    stacked [conditional | unconditional], equal scalars pass through, anything that does not line up row for row disables the
    merge (the loop then runs two forwards)."""
    import types
    import numpy as np
    from xdiffusion_b200.diffusion.ddpm import _DeviceLoop
    B = 3
    cond = {"text_embeddings": torch.randn(B, 5, 8), "classes": torch.arange(B), "flag": 7}
    unc = {"text_embeddings": torch.zeros(B, 5, 8), "classes": torch.full((B,), 10), "flag": 7}
    both = _DeviceLoop._merge(types.SimpleNamespace(context=cond, uncond=unc), B)
    assert both["flag"] == 7 and both["classes"].tolist() == [0, 1, 2, 10, 10, 10]
    assert torch.equal(both["text_embeddings"][:B], cond["text_embeddings"]) and float(both["text_embeddings"][B:].abs().max()) == 0
    assert _DeviceLoop._merge(types.SimpleNamespace(context=cond, uncond={**unc, "flag": 8}), B) is None
    assert _DeviceLoop._merge(types.SimpleNamespace(context=cond, uncond={k: v for k, v in unc.items() if k != "flag"}), B) is None
    assert _DeviceLoop._merge(types.SimpleNamespace(context=cond, uncond={**unc, "classes": torch.zeros(B + 1)}), B) is None
    arr = {"a": np.zeros(2)}
    assert _DeviceLoop._merge(types.SimpleNamespace(context={**cond, **arr}, uncond={**unc, **arr}), B) is None

"""The oracle (CPU restatement) against fixtures produced by the real reference
(tests/golden/make_golden.py).  Integer/table work must be bit-exact; fp32 network
outputs differ only by summation order (einsum vs matmul) -> 1e-5 relative L2."""
import os

import pytest
import torch

from oracle import nets, samplers, schedules
from oracle.loop import OracleModel
from oracle.weights import synth_state_dict

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def have(name):
    return os.path.exists(os.path.join(GOLDEN, f"{name}.pt"))


def test_discrete_tables_bit_exact(golden):
    kat = golden("kat")
    t = schedules.discrete_tables(1000, "linear")
    for k, v in kat["discrete"].items():
        assert torch.equal(t[k], v), k
    assert torch.equal(schedules.fixed_large_logvar(t), kat["fixed_large_logvar"])
    # SURVEY.md S4 known answers
    assert abs(float(t["sqrt_recip_alphas_cumprod"][500]) - 3.5852506161) < 1e-6
    assert abs(float(kat["fixed_large_logvar"][0]) + 9.8167247772) < 1e-6


def test_continuous_tables_and_index_rule_bit_exact(golden):
    kat = golden("kat")
    g = schedules.cosine_logsnr_table(1024, -20, 20)
    assert torch.equal(g, kat["gammas"])
    for N in (1024, 1000, 50):
        ls = torch.stack([g[schedules.continuous_indices(i, N, 1024)[0]] for i in range(N)])
        lt = torch.stack([g[schedules.continuous_indices(i, N, 1024)[1]] for i in range(N)])
        assert torch.equal(ls, kat[f"logsnr_s_{N}"]) and torch.equal(lt, kat[f"logsnr_t_{N}"])


def test_embeddings_bit_exact(golden):
    kat = golden("kat")
    tt = torch.tensor([0, 1, 500, 999])
    assert torch.equal(nets.sinusoid_unet(tt, 128, 1000.0), kat["sin_unet_1000"])
    assert torch.equal(nets.sinusoid_unet(torch.tensor([0.001, 0.5, 0.999]), 128, 1.0), kat["sin_unet_1"])
    assert torch.equal(nets.sinusoid_dit(tt, 256), kat["sin_dit"])
    assert torch.equal(nets.sincos_pos_embed_2d(384, 4, 16), kat["pos_dit"])
    assert torch.equal(nets.sincos_pos_embed_2d(384, 4, 4), kat["pos_pixart"])


def test_dynamic_threshold_bit_exact(golden):
    kat = golden("kat")
    assert torch.equal(samplers.dynamic_threshold(kat["dyn_in"], 0.99, 1.7), kat["dyn_out"])


def _model(fx):
    sd = synth_state_dict(fx["manifest"], seed=0)
    kind = fx["kind"]
    if "pos_embed" in fx["manifest"]:
        base = 16 if kind == "dit" else 4
        sd["pos_embed"] = nets.sincos_pos_embed_2d(384, 4, base)[None]
    return OracleModel(kind, fx["config"], sd)


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5", "c6", "c7"])
def test_single_steps(name, golden):
    if not have(name):
        pytest.skip("fixture not generated")
    fx = golden(name)
    m = _model(fx)
    N = m.steps
    for i, s in fx["steps"].items():
        B = s["x"].shape[0]
        ctx = dict(fx["ctx"])
        if m.sched_kind == "ContinuousNoiseScheduler":
            idx_s, idx_t = schedules.continuous_indices(i, N, N)
            assert torch.equal(m.gammas[idx_s], s["logsnr_s"]) and torch.equal(m.gammas[idx_t], s["logsnr_t"])
            ctx["logsnr_t"] = m.gammas[idx_t].expand(B)
            t = torch.full((B,), schedules.continuous_time(i, N))
        elif m.sched_kind == "DiscreteNoiseScheduler":
            t = torch.full((B,), i, dtype=torch.int64)
        else:
            t = torch.full((B,), schedules.rectified_flow_time(i))
        o = m.score(s["x"], t, ctx)
        assert rel_l2(o, s["score"]) < 1e-5, (name, i, rel_l2(o, s["score"]))
        # the sampler update itself, fed the reference's own score: bit-exact
        if m.sched_kind == "DiscreteNoiseScheduler":
            xn = samplers.ancestral_discrete(s["x"], s["score"], s["z"], i, m.tables, m.logvar,
                                             m.prediction, m.threshold)
        elif m.sched_kind == "ContinuousNoiseScheduler":
            xn = samplers.ancestral_continuous(s["x"], s["score"], s["z"], i, s["logsnr_s"], s["logsnr_t"],
                                               m.prediction, m.threshold)
            if "x_next_ddim" in s:
                xd = samplers.ddim_continuous(s["x"], s["score"], i, s["logsnr_s"], s["logsnr_t"],
                                              m.prediction, m.threshold)
                assert torch.equal(xd, s["x_next_ddim"]), (name, i, "ddim")
        else:
            xn = samplers.euler_flow(s["x"], s["score"], m.N)
        assert torch.equal(xn, s["x_next"]), (name, i, float((xn - s["x_next"]).abs().max()))


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5", "c6", "c7"])
def test_short_loop(name, golden):
    if not have(name):
        pytest.skip("fixture not generated")
    fx = golden(name)
    m = _model(fx)
    lp = fx["loop"]
    out = m.sample(lp["x_T"], lp["noises"], ctx=fx["ctx"], num_sampling_steps=lp["K"])
    assert rel_l2(out, lp["samples"]) < 1e-4
    if "loop_cfg" in fx:
        if m.kind == "dit":
            un = {"classes": torch.full_like(fx["ctx"]["classes"], 10)}
        else:
            B = lp["x_T"].shape[0]
            un = {"text_embeddings": fx["null_embedding"][None].expand(B, -1, -1)}
        out = m.sample(lp["x_T"], lp["noises"], ctx=fx["ctx"], num_sampling_steps=lp["K"],
                       cfg_scale=fx["loop_cfg"]["w"], uncond_ctx=un)
        assert rel_l2(out, fx["loop_cfg"]["samples"]) < 1e-4

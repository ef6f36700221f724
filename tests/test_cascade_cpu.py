"""SURVEY 8(f1), second half, host side: the oracle's efficient UNet + super-resolution input pipeline + stage chaining
against the REAL reference (tests/golden/make_cascade.py), and state-dict key compatibility of the product's SR stage."""
import torch

from oracle import nets, samplers
from tests.helpers import oracle_model, rel_l2


def test_oracle_sr_stage_single_steps(golden):
    fx = golden("c8")
    om = oracle_model(fx)
    for i, s in fx["steps"].items():
        B = s["x"].shape[0]
        t = torch.full((B,), i, dtype=torch.int64)
        ctx = dict(fx["ctx"], z_cond=s["z_cond"])
        sr_t = (torch.ones(B, dtype=torch.long) * om.steps * om.sr["level"]).to(torch.long)
        xin = nets.sr_input(s["x"], ctx["low_resolution_images"], 32, om.tables, sr_t, s["z_cond"])
        assert rel_l2(xin, s["xin"]) < 1e-6
        o = om.score(s["x"], t, ctx)
        assert rel_l2(o, s["score"]) < 1e-5, (i, rel_l2(o, s["score"]))
        xn = samplers.ancestral_discrete(s["x"], s["score"], s["z"], i, om.tables, om.logvar, om.prediction, om.threshold)
        assert torch.equal(xn, s["x_next"]), i


def test_oracle_sr_loop_and_cascade(golden):
    fx = golden("c8")
    om = oracle_model(fx)
    lp = fx["loop"]
    out = om.sample(lp["x_T"], lp["noises"], ctx=fx["ctx"], num_sampling_steps=lp["K"], cond_noises=lp["cond_noises"])
    assert rel_l2(out, lp["samples"]) < 1e-4
    # the cascade: stage 1 (c7 network) -> its [0, 1] samples condition stage 2 (diffusion/cascade.py:148-179)
    cs = fx["cascade"]
    base = oracle_model(golden("c7"))
    text = {"text_embeddings": fx["ctx"]["text_embeddings"]}
    s1 = base.sample(cs["x_T"][0], cs["noises"][0], ctx=text, num_sampling_steps=cs["K"])
    assert rel_l2(s1, cs["stage_samples"][0]) < 1e-4
    s2 = om.sample(cs["x_T"][1], cs["noises"][1], ctx=dict(text, low_resolution_images=s1), num_sampling_steps=cs["K"],
                   cond_noises=cs["cond_noises"])
    assert rel_l2(s2, cs["stage_samples"][1]) < 1e-4

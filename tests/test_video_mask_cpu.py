"""SURVEY 8(f4), host side: the frame-index scheme reproduces the reference's sequences (fixture written by the real
``Autoregressive``), and the oracle's masked loop reproduces the reference's conditional video samples."""
import os

import torch

from tests.conftest import GOLDEN
from tests.helpers import oracle_model, rel_l2
from xdiffusion_b200.samplers.schemes import Autoregressive


def _fixture():
    return torch.load(os.path.join(GOLDEN, "mask_c5.pt"), weights_only=False)


def test_autoregressive_scheme_matches_reference_sequences():
    for rec in _fixture()["schemes"]:
        length, max_frames, step = rec["args"]
        sch = Autoregressive(video_length=length, num_observed_frames=0, max_frames=max_frames, step_size=step)
        it = iter(sch)
        it.set_videos([0, 1])
        got = []
        while True:
            try:
                got.append(next(it))
            except StopIteration:
                break
        assert len(got) == len(rec["seq"])
        for (o, l, m), (ro, rl, rm) in zip(got, rec["seq"]):
            assert o == ro and l == rl and torch.equal(m, rm)
        assert sch.is_done() and sch.video_length == length and sch.num_observations == 0


def test_oracle_masked_loop_matches_reference(golden):
    lp = _fixture()["loop"]
    om = oracle_model(golden("c5"))
    out = om.sample(lp["x_T"], lp["noises"], ctx={"video_mask": lp["video_mask"], "x0": lp["x0"]},
                    num_sampling_steps=lp["K"])
    assert rel_l2(out, lp["samples"]) < 1e-4
    keep = ~lp["video_mask"][0]
    assert torch.equal(out[:, :, keep], ((lp["x0"].clamp(-1, 1) + 1) * 0.5)[:, :, keep])

"""Video-mask fixture: the REAL reference's conditional video sampling and frame-index scheme.

    python tests/golden/make_mask.py            # writes tests/golden/mask_c5.pt

* ``loop``: C5 (video UNet-3D) ``sample()`` for K = 4 steps with ``context["video_mask"]`` (first 6 of 16 frames observed)
  and ``context["x0"]``: the reference blends x_t with x0 before and after every step (diffusion/ddpm.py:963-982).
* ``schemes``: the (observed, latent, mask) sequences of the reference's ``Autoregressive`` scheme (samplers/schemes.py)
  for a few (video_length, max_frames, step_size) settings.
Only runs in the authoring container.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.golden.make_golden import NoiseFeeder, build  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.manual_seed(1234)
    model, kind, manifest = build("c5")
    from xdiffusion.samplers.schemes import Autoregressive
    g = torch.Generator().manual_seed(2468)
    B, K, shape = 1, 4, (1, 1, 16, 32, 32)
    x_T = torch.randn(shape, generator=g)
    zs = [torch.randn(shape, generator=g) for _ in range(K)]
    x0 = torch.randn(shape, generator=g).clamp(-1, 1)
    mask = torch.ones(B, 16, dtype=torch.bool)
    mask[:, :6] = False
    with NoiseFeeder() as feeder:
        feeder.queue = [zs[i].clone() for i in reversed(range(K))]
        samples, _ = model.sample(context={"video_mask": mask, "x0": x0}, num_samples=B, num_sampling_steps=K,
                                  initial_noise=x_T.clone())
    out = {"loop": {"K": K, "x_T": x_T, "noises": zs, "x0": x0, "video_mask": mask, "samples": samples}, "schemes": []}
    for (length, max_frames, step) in [(40, 16, 8), (30, 16, 4), (16, 16, 8), (21, 10, 3)]:
        sch = Autoregressive(video_length=length, num_observed_frames=0, max_frames=max_frames, step_size=step)
        it = iter(sch)
        it.set_videos([0, 1])
        seq = []
        while True:
            try:
                o, l, m = next(it)
            except StopIteration:
                break
            seq.append((o, l, m.clone()))
        out["schemes"].append({"args": (length, max_frames, step), "seq": seq})
    torch.save(out, os.path.join(HERE, "mask_c5.pt"))
    print("mask_c5 written; observed frames kept:",
          bool(torch.equal(samples[:, :, :6], ((x0.clamp(-1, 1) + 1) * 0.5)[:, :, :6])),
          [len(s["seq"]) for s in out["schemes"]])


if __name__ == "__main__":
    main()

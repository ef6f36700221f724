"""Cascade fixture (second half): the REAL reference's super-resolution stage and the two-stage Imagen cascade.

    python tests/golden/make_cascade.py            # writes tests/golden/c8.pt

``configs/image/mnist/imagen_8x8_to_32x32.yaml`` (efficient UNet, low-resolution conditioning concatenated to the input after a
Gaussian conditioning augmentation that is RE-DRAWN at every network evaluation, augmentation level added to the timestep
embedding) with the T5 pieces stripped and synthetic text embeddings, like c7.  Stored: single steps (score, x_next), a K = 4
loop, and the cascade chain of ``diffusion/cascade.py:148-179`` (stage 1 = c7's model for K = 4 steps, its samples conditioning
stage 2 for K = 4 steps).  Noise order per reverse step: the conditioning augmentation first (InputPreprocessor, inside
``process_input``), then the sampler's z.  Only runs in the authoring container.
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
    sr, kind, manifest = build("c8")
    base, _, _ = build("c7")
    cfg = sr.config().to_dict()
    g = torch.Generator().manual_seed(8642)
    B, shape = 2, (2, 1, 32, 32)
    text = torch.randn(B, 77, 768, generator=g)
    low = torch.rand(B, 1, 8, 8, generator=g)                       # low-resolution images in [0, 1]
    level = cfg["super_resolution"]["sampling_augmentation_level"]
    out = {"manifest": manifest, "config": cfg, "kind": kind, "ctx": {"text_embeddings": text, "low_resolution_images": low}}
    sampler = sr._reverse_process_sampler
    steps = {}
    with NoiseFeeder() as feeder, torch.no_grad():
        for i in [999, 500, 1, 0]:
            x = torch.randn(shape, generator=g)
            zc = torch.randn(shape, generator=g)
            z = torch.randn(shape, generator=g)
            c = {"text_embeddings": text, "text_prompts": ["", ""], "low_resolution_images": low, "augmentation_level": level,
                 "timestep": torch.tensor([i] * B), "timestep_idx": i}
            feeder.queue = [zc.clone()]
            c1 = dict(c)
            xin = sr.process_input(x=x, context=c1)
            score = sr.predict_score(xin, context=c1)
            feeder.queue = [zc.clone(), z.clone()]
            x_next = sampler.p_sample(x, context=dict(c), unconditional_context=None, diffusion_model=sr)
            assert not feeder.queue
            steps[i] = {"x": x, "z_cond": zc, "z": z, "xin": xin, "score": score, "x_next": x_next}
    out["steps"] = steps
    K = 4
    x_T = torch.randn(shape, generator=g)
    zcs = [torch.randn(shape, generator=g) for _ in range(K)]
    zs = [torch.randn(shape, generator=g) for _ in range(K)]
    ctx = {"text_embeddings": text, "text_prompts": ["", ""], "low_resolution_images": low}
    with NoiseFeeder() as feeder:
        feeder.queue = [t.clone() for i in reversed(range(K)) for t in (zcs[i], zs[i])]
        samples, _ = sr.sample(context=dict(ctx), num_samples=B, num_sampling_steps=K, initial_noise=x_T.clone())
        assert not feeder.queue
    out["loop"] = {"K": K, "x_T": x_T, "cond_noises": zcs, "noises": zs, "samples": samples}
    # ---- the cascade (diffusion/cascade.py:148-179 with K steps per stage): stage 1 = imagen_base (c7 weights)
    shape1 = (B, 1, 8, 8)
    x1 = torch.randn(shape1, generator=g)
    z1 = [torch.randn(shape1, generator=g) for _ in range(K)]
    x2 = torch.randn(shape, generator=g)
    zc2 = [torch.randn(shape, generator=g) for _ in range(K)]
    z2 = [torch.randn(shape, generator=g) for _ in range(K)]
    stage_out = []
    prev = None
    for model, xT, queue in ((base, x1, [z1[i] for i in reversed(range(K))]),
                             (sr, x2, [t for i in reversed(range(K)) for t in (zc2[i], z2[i])])):
        c = {"text_embeddings": text, "text_prompts": ["", ""]}
        if prev is not None:
            c[model.config().super_resolution.conditioning_key] = prev
        with NoiseFeeder() as feeder:
            feeder.queue = [t.clone() for t in queue]
            prev, _ = model.sample(context=c, num_samples=B, num_sampling_steps=K, initial_noise=xT.clone())
            assert not feeder.queue
        stage_out.append(prev)
    out["cascade"] = {"K": K, "x_T": [x1, x2], "noises": [z1, z2], "cond_noises": zc2, "stage_samples": stage_out}
    torch.save(out, os.path.join(HERE, "c8.pt"))
    print("c8 written", {i: float(s["score"].abs().mean()) for i, s in steps.items()}, float(stage_out[1].mean()))


if __name__ == "__main__":
    main()

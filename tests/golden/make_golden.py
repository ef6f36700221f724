"""Generate golden fixtures by running the REAL reference (imported from
/root/reference) on seeded synthetic weights, inputs and injected noise.

    python tests/golden/make_golden.py            # writes tests/golden/c{1..5}.pt, kat.pt

Only runs in the authoring container (the reference tree does not travel).  The
fixtures are small: the weights are regenerated from the manifest by
``oracle.weights.synth_state_dict``; only inputs/outputs are stored.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_bootstrap  # noqa: E402
from oracle.weights import canonical_manifest, synth_state_dict  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    "c1": ("configs/image/mnist/ddpm_32x32_epsilon_discrete.yaml", "unet"),
    "c2": ("configs/image/mnist/dit.yaml", "dit"),
    "c3": ("configs/image/mnist/rectified_flow_32x32.yaml", "unet"),
    "c4": ("configs/image/mnist/pixart_alpha.yaml", "pixart"),
    "c5": ("configs/video/moving_mnist/video_diffusion_models.yaml", "unet3d"),
    # DDIM needs the continuous scheduler (samplers/ddim.py:43-45)
    "c6": ("configs/image/mnist/ddpm_32x32_v_continuous.yaml", "unet"),
    # text-conditioned UNet (Imagen base stage, 8x8): encoder K/V in front of the self-attention K/V,
    # attention-pooled text embedding added to the timestep embedding; synthetic T5-shaped embeddings
    "c7": ("configs/image/mnist/imagen_base.yaml", "unet"),
    # ... and its 8x8 -> 32x32 super-resolution stage (efficient UNet; fixture written by make_cascade.py)
    "c8": ("configs/image/mnist/imagen_8x8_to_32x32.yaml", "effunet"),
}
PATCH = {"c4": ref_bootstrap.strip_t5_from_pixart, "c7": ref_bootstrap.strip_t5_from_imagen,
         "c8": ref_bootstrap.strip_t5_from_imagen}


class NoiseFeeder:
    """Replaces torch.randn_like inside the reference samplers with a queue."""

    def __init__(self):
        self.queue = []
        self._orig = torch.randn_like

    def __enter__(self):
        feeder = self

        def fake(x, *a, **k):
            if feeder.queue:
                z = feeder.queue.pop(0)
                assert z.shape == x.shape
                return z
            return feeder._orig(x, *a, **k)

        torch.randn_like = fake
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig


def patch_cfg_bug():
    """ancestral.py:221-227 calls _pred_epsilon without diffusion_model (TypeError).
    Supply it from the enclosing call, changing nothing else."""
    from xdiffusion.samplers import ancestral
    cls = ancestral.AncestralSampler
    if getattr(cls, "_xdb_patched", False):
        return
    orig_hat, orig_eps = cls._pred_x_hat, cls._pred_epsilon

    def pred_x_hat(self, z_t, context, unconditional_context, diffusion_model, **kw):
        self._xdb_dm = diffusion_model
        return orig_hat(self, z_t=z_t, context=context, unconditional_context=unconditional_context,
                        diffusion_model=diffusion_model, **kw)

    def pred_eps(self, x, context, diffusion_model=None, epsilon_v_param=None):
        return orig_eps(self, x=x, context=context,
                        diffusion_model=diffusion_model or self._xdb_dm, epsilon_v_param=epsilon_v_param)

    cls._pred_x_hat, cls._pred_epsilon, cls._xdb_patched = pred_x_hat, pred_eps, True


def build(name):
    rel, kind = CONFIGS[name]
    model = ref_bootstrap.load_reference_model(rel, PATCH.get(name))
    manifest = canonical_manifest(model.state_dict())
    sd = synth_state_dict(manifest, seed=0)
    missing, unexpected = model.load_state_dict({"_score_network." + k: v for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    return model, kind, manifest


def make(name):
    torch.manual_seed(1234)
    model, kind, manifest = build(name)
    cfg = model.config().to_dict()
    B = 1 if kind == "unet3d" else 2
    size = cfg["diffusion"]["sampling"]["output_spatial_size"]
    shape = (B, 1, 16, size, size) if kind == "unet3d" else (B, 1, size, size)
    g = torch.Generator().manual_seed(4321)
    out = {"manifest": manifest, "config": cfg, "kind": kind}
    ctx = {}
    if kind == "dit":
        ctx["classes"] = torch.tensor([3, 7])
    if kind == "pixart":
        ctx["classes"] = torch.tensor([3, 7])
        ctx["text_prompts"] = ["", ""]
        ctx["text_embeddings"] = torch.randn(B, 77, 768, generator=g)
    if name == "c7":
        ctx["text_prompts"] = ["", ""]
        ctx["text_embeddings"] = torch.randn(B, 77, 768, generator=g)
    out["ctx"] = {k: v for k, v in ctx.items() if torch.is_tensor(v)}
    sched = model.noise_scheduler()
    continuous = sched.continuous()
    N = sched.steps()
    sampler = model._reverse_process_sampler

    # ---- single steps at chosen loop indices through the reference sampler -------------
    steps = {}
    idxs = [N - 1, N // 2, 1, 0]
    with NoiseFeeder() as feeder, torch.no_grad():
        for i in idxs:
            x = torch.randn(shape, generator=g)
            z = torch.randn(shape, generator=g)
            c = dict(ctx)
            t = torch.tensor([i] * B)
            c["timestep_idx"] = i
            if continuous:
                c["logsnr_s"] = sched.logsnr(t / N)
                c["logsnr_t"] = sched.logsnr((t + 1) / N)
                t = t / N
            c["timestep"] = t
            c_net = dict(c)
            if name == "c3":
                from oracle.schedules import rectified_flow_time
                c_net["timestep"] = torch.ones(B) * rectified_flow_time(i)
            score = model.predict_score(x, context=dict(c_net))
            feeder.queue = [z.clone()]
            x_next = sampler.p_sample(x, context=dict(c), unconditional_context=None, diffusion_model=model)
            feeder.queue = []
            steps[i] = {"x": x, "z": z, "score": score, "x_next": x_next}
            if continuous:
                steps[i]["logsnr_s"], steps[i]["logsnr_t"] = c["logsnr_s"][0].clone(), c["logsnr_t"][0].clone()
            if name == "c6":
                from xdiffusion.samplers.ddim import DDIMSampler
                steps[i]["x_next_ddim"] = DDIMSampler().p_sample(
                    x, context=dict(c), unconditional_context=None, diffusion_model=model)
    out["steps"] = steps

    # ---- a short full loop through sample() ----------------------------------------------
    K = 4
    x_T = torch.randn(shape, generator=g)
    zs = [torch.randn(shape, generator=g) for _ in range(K)]
    with NoiseFeeder() as feeder:
        # the loop runs i = K-1 .. 0 and draws one randn_like per step, in that order
        feeder.queue = [zs[i].clone() for i in reversed(range(K))]
        samples, _ = model.sample(context=dict(ctx), num_samples=B, num_sampling_steps=K,
                                  initial_noise=x_T.clone())
    out["loop"] = {"K": K, "x_T": x_T, "noises": zs, "samples": samples}

    # ---- classifier-free guidance (C2 classes, C4 text embeddings) -------------------------
    if kind in ("dit", "pixart") or name == "c7":
        patch_cfg_bug()
        w = 2.0
        if kind == "pixart" or name == "c7":
            from xdiffusion.context import UnconditionalEmbeddingAdapter
            adapter = UnconditionalEmbeddingAdapter([77, 768])
            with torch.no_grad():
                adapter.y_embedding.copy_(torch.randn(77, 768, generator=g) / 768 ** 0.5)
            model._unconditional_context = adapter
            out["null_embedding"] = adapter.y_embedding.detach().clone()
        with NoiseFeeder() as feeder:
            feeder.queue = [zs[i].clone() for i in reversed(range(K))]
            samples, _ = model.sample(context=dict(ctx), num_samples=B, num_sampling_steps=K,
                                      initial_noise=x_T.clone(), classifier_free_guidance=w)
        out["loop_cfg"] = {"w": w, "samples": samples}
    torch.save(out, os.path.join(HERE, f"{name}.pt"))
    print(name, "written", {i: float(s["score"].abs().mean()) for i, s in steps.items()})


def make_kat():
    """Known-answer vectors for tables and embeddings."""
    ref_bootstrap.bootstrap()
    from xdiffusion.layers.embedding import SinusoidalPositionEmbedding
    from xdiffusion.layers.utils import get_2d_sincos_pos_embed, timestep_embedding
    from xdiffusion.scheduler import ContinuousNoiseScheduler, DiscreteNoiseScheduler
    from xdiffusion.utils import dynamic_thresholding
    imp = {"target": "xdiffusion.importance_sampling.UniformSampler", "params": {"num_timesteps": 1000}}
    d = DiscreteNoiseScheduler("linear", 1000, "l2", importance_sampler=imp)
    kat = {"discrete": {k: v.clone() for k, v in d.state_dict().items()}}
    ctx = {"timestep": torch.arange(1000)}
    kat["fixed_large_logvar"] = d.variance_fixed_large(ctx, (1000,))[1].clone()
    c = ContinuousNoiseScheduler(1024, "cosine", "l2", -20, 20)
    kat["gammas"] = c.gammas.clone()
    for N in (1024, 1000, 50):
        t = torch.arange(N)
        kat[f"logsnr_s_{N}"] = c.logsnr(t / N).clone()
        kat[f"logsnr_t_{N}"] = c.logsnr((t + 1) / N).clone()
    tt = torch.tensor([0, 1, 500, 999])
    kat["sin_unet_1000"] = SinusoidalPositionEmbedding(128, max_time=1000.0)(tt)
    kat["sin_unet_1"] = SinusoidalPositionEmbedding(128, max_time=1.0)(torch.tensor([0.001, 0.5, 0.999]))
    kat["sin_dit"] = timestep_embedding(tt, 256)
    kat["pos_dit"] = torch.from_numpy(get_2d_sincos_pos_embed(384, 4)).float()
    kat["pos_pixart"] = torch.from_numpy(get_2d_sincos_pos_embed(384, 4, lewei_scale=(1.0,), base_size=4)).float()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 1, 32, 32, generator=g) * torch.tensor([0.3, 1.0, 2.0, 5.0])[:, None, None, None]
    kat["dyn_in"], kat["dyn_out"] = x, dynamic_thresholding(x, p=0.99, c=1.7)
    torch.save(kat, os.path.join(HERE, "kat.pt"))
    print("kat written")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS) + ["kat"]
    for n in names:
        make_kat() if n == "kat" else make(n)

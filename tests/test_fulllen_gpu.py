"""Parity at the BASELINE configurations and at full length (GPU).

* ``test_full_length_matches_reference``: C1 (batch 64), C2 (batch 1024) and C3 (batch 64) run their COMPLETE
  reverse process (1000 timesteps) on the product path at the benchmark batch size; the first rows carry the x_T /
  per-step noise / classes of ``tests/golden/full_<name>.pt``, which holds the samples the REAL reference produced
  for those rows (tests/golden/make_fulllen.py).  Rows are independent on this path, so they must agree -- this is
  the "final samples within a stated tolerance" bar of BASELINE.json.
* ``test_medium_loop_*``: >= 50-step loops of C3, C4 (with classifier-free guidance), C5 and C6 (ancestral and DDIM)
  against the CPU oracle on the same inputs.
* ``test_c5_unit_gain_error_is_bounded``: the video network with its temporal qkv projections at unit gain (the
  regime the softened fixture avoids): the measured error is asserted, not hidden.
* ``test_reloading_weights_invalidates_the_captured_loop``: ADVICE r1 (high).

FINAL-SAMPLE TOLERANCE (stated here and in DESIGN.md): relative L2 of the final [0,1] images over the compared rows.
The bf16 tensor-core path perturbs every score by <= 2e-2 relative; the reverse process is contractive in the
injected-noise setting (x0 clipping / thresholding, posterior mean pulls towards x0), so the error does not grow with
the step count: measured values are recorded next to each bound.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.conftest import GOLDEN  # noqa: E402
from tests.golden.make_fulllen import fulllen_inputs  # noqa: E402
from tests.helpers import fixture_state_dict, oracle_model, product_model, rel_l2  # noqa: E402

DEV = "cuda"
#: final-sample tolerance after the full 1000-step loop (relative L2 on [0,1] images, bf16 path vs fp32 reference).
#: Measured on B200 (round 2, first GPU run): c2 1.4e-3 (worst row 1.8e-3), c1 2.1e-3 (2.3e-3), c3 1.7e-3 (1.8e-3).
FULL_TOL = {"c1": 1e-2, "c2": 1e-2, "c3": 1e-2}
BENCH_BATCH = {"c1": 64, "c2": 1024, "c3": 64}
MEDIUM_TOL = 2e-2


@pytest.mark.parametrize("name", ["c2", "c1", "c3"])
def test_full_length_matches_reference(name, golden):
    from xdiffusion_b200 import ops
    full = torch.load(os.path.join(GOLDEN, f"full_{name}.pt"), weights_only=False)
    rows, N, B = full["rows"], full["steps"], BENCH_BATCH[name]
    x_T, noise, classes = fulllen_inputs(name, rows, full["seed"], N)
    assert torch.equal(x_T.flatten()[:8], full["probe"]["x_T"]) and torch.equal(noise.flatten()[-8:], full["probe"]["noise"])
    assert full["oracle_vs_reference"] < 1e-3            # the CPU restatement tracks the reference over the whole loop
    m = product_model(golden(name))
    g = torch.Generator(device=DEV).manual_seed(77)
    x_all = torch.randn(B, 1, 32, 32, device=DEV, generator=g)
    z_all = torch.randn(N, B, 1, 32, 32, device=DEV, generator=g)
    x_all[:rows] = x_T.to(DEV)
    z_all[:, :rows] = noise.to(DEV)
    ctx = {}
    if classes is not None:
        c_all = torch.randint(0, 10, (B,), device=DEV, generator=g)
        c_all[:rows] = classes.to(DEV)
        ctx = {"classes": c_all}
    out, _ = m.sample(context=dict(ctx), num_samples=B, num_sampling_steps=N, initial_noise=x_all, noise=z_all)
    err = rel_l2(out[:rows], full["samples"])
    worst = max(rel_l2(out[r:r + 1], full["samples"][r:r + 1]) for r in range(rows))
    print(f"[full-length] {name}: B={B} N={N} rel L2 over {rows} rows = {err:.3e} (worst row {worst:.3e})")
    assert err < FULL_TOL[name], (name, err)
    assert bool(torch.isfinite(out).all()) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    # the same rows alone, bit for bit, without split-K (shard independence at full length)
    ops.set_split_k(False)
    try:
        a, _ = m.sample(context=dict(ctx), num_samples=B, num_sampling_steps=N, initial_noise=x_all, noise=z_all)
        sub_ctx = {k: v[:rows].contiguous() for k, v in ctx.items()}
        b, _ = m.sample(context=sub_ctx, num_samples=rows, num_sampling_steps=N, initial_noise=x_all[:rows].contiguous(),
                        noise=z_all[:, :rows].contiguous())
        assert torch.equal(a[:rows], b)
    finally:
        ops.set_split_k(True)


def _medium(name, golden, B, K, sampler=None, cfg=None):
    fx = golden(name)
    m, om = product_model(fx), oracle_model(fx)
    g = torch.Generator().manual_seed(31)
    size = fx["config"]["diffusion"]["sampling"]["output_spatial_size"]
    shape = (B, 1, 16, size, size) if fx["kind"] == "unet3d" else (B, 1, size, size)
    x_T = torch.randn(shape, generator=g)
    noise = torch.randn((K,) + shape, generator=g)
    ctx, uncond = {}, None
    if "null_embedding" in fx:                           # text-conditioned: c4 (PixArt), c7 (Imagen-base UNet)
        ctx["text_embeddings"] = torch.randn(B, 77, 768, generator=g)
        if cfg is not None:
            from xdiffusion_b200.context import UnconditionalEmbeddingAdapter
            ad = UnconditionalEmbeddingAdapter([77, 768])
            ad.y_embedding.copy_(fx["null_embedding"])
            m._unconditional_context = ad.to(DEV)
            uncond = {"text_embeddings": fx["null_embedding"][None].expand(B, -1, -1).contiguous()}
    ref = om.sample(x_T, [noise[i] for i in range(K)], ctx=dict(ctx), num_sampling_steps=K,
                    sampler=sampler or "ancestral", cfg_scale=cfg, uncond_ctx=uncond)
    kw = {}
    if sampler == "ddim":
        from xdiffusion_b200.samplers.ddim import DDIMSampler
        kw["sampler"] = DDIMSampler()
    if cfg is not None:
        kw["classifier_free_guidance"] = cfg
    pctx = {k: v.to(DEV) for k, v in ctx.items()}
    if fx["kind"] == "pixart":          # read by RunProjection("classes") but dropped by the combine (drop_prob = 1)
        pctx["classes"] = torch.zeros(B, dtype=torch.long, device=DEV)
    out, _ = m.sample(context=pctx, num_samples=B, num_sampling_steps=K, initial_noise=x_T.to(DEV),
                      noise=noise.to(DEV), **kw)
    err = rel_l2(out, ref)
    print(f"[medium loop] {name} sampler={sampler or 'default'} cfg={cfg}: B={B} K={K} rel L2 = {err:.3e}")
    return err


def test_medium_loop_c3_rectified_flow(golden):
    assert _medium("c3", golden, 4, 100) < MEDIUM_TOL


def test_medium_loop_c4_pixart_cfg(golden):
    assert _medium("c4", golden, 4, 50) < MEDIUM_TOL
    assert _medium("c4", golden, 4, 50, cfg=2.0) < MEDIUM_TOL


def test_medium_loop_c7_text_conditioned_unet_cfg(golden):
    """Imagen-base UNet (8x8) with encoder K/V in front of the self-attention K/V and the attention-pooled
    text embedding in the timestep embedding, classifier-free guidance w = 2, 50 steps with dynamic thresholding."""
    assert _medium("c7", golden, B=8, K=50, cfg=2.0) < MEDIUM_TOL


def test_medium_loop_c5_video(golden):
    assert _medium("c5", golden, 1, 50) < MEDIUM_TOL


@pytest.mark.parametrize("sampler", ["ancestral", "ddim"])
def test_medium_loop_c6_continuous(golden, sampler):
    assert _medium("c6", golden, 4, 64, sampler=sampler) < MEDIUM_TOL


def test_c5_unit_gain_error_is_bounded(golden):
    """The committed C5 fixture scales the temporal qkv projections by 0.35 (oracle/weights.py) because the reference's
    UNSCALED temporal logits make unit-gain random weights a one-hot softmax that amplifies any rounding.  This test
    runs the configuration as the reference initialises it (gain 1.0) and asserts the measured score error
    (3.8e-2 in round 1) stays under 6e-2 -- looser than the 2e-2 bar, recorded as such in DESIGN.md."""
    from oracle import weights as oweights
    from oracle.loop import OracleModel
    from xdiffusion_b200.diffusion import GaussianDiffusion_DDPM
    from xdiffusion_b200.utils import DotConfig
    fx = golden("c5")
    old = oweights.TEMPORAL_QKV_GAIN
    oweights.TEMPORAL_QKV_GAIN = 1.0
    try:
        sd = fixture_state_dict(fx)
    finally:
        oweights.TEMPORAL_QKV_GAIN = old
    om = OracleModel(fx["kind"], fx["config"], sd)
    m = GaussianDiffusion_DDPM(DotConfig(fx["config"]))
    m.load_state_dict({"_score_network." + k: v for k, v in sd.items()}, strict=False)
    m = m.to(DEV).eval()
    i = 512
    s = fx["steps"][i]
    x = s["x"]
    lam_t = s["logsnr_t"]
    ref = om.score(x, None, {"logsnr_t": lam_t.expand(x.shape[0])})
    got = m.predict_score(x.to(DEV), context={"logsnr_t": lam_t.expand(x.shape[0]).to(DEV),
                                              "timestep": torch.full((x.shape[0],), i / 1024.0, device=DEV)})
    err = rel_l2(got, ref)
    print(f"[c5 unit gain] score rel L2 = {err:.3e}")
    assert err < 6e-2, err


def test_reloading_weights_invalidates_the_captured_loop(golden):
    """sample(); load different weights into the SAME model object; sample() again with the same shape / step count /
    context signature: the second call must not replay the loop captured with the old repacked weights."""
    fx = golden("c2")
    m = product_model(fx)
    g = torch.Generator().manual_seed(2)
    B, K = 8, 5
    x_T = torch.randn(B, 1, 32, 32, generator=g).to(DEV)
    ctx = {"classes": torch.randint(0, 10, (B,), generator=g).to(DEV)}
    a, _ = m.sample(context=dict(ctx), num_samples=B, num_sampling_steps=K, initial_noise=x_T, seed=5)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g2 = torch.Generator().manual_seed(99)
    new_sd = {k: (v + 0.05 * torch.randn(v.shape, generator=g2).to(v.device) * v.abs().mean()
                  if v.is_floating_point() and k.startswith("_score_network.blocks") else v) for k, v in sd.items()}
    m.load_state_dict(new_sd)
    b, _ = m.sample(context=dict(ctx), num_samples=B, num_sampling_steps=K, initial_noise=x_T, seed=5)
    fresh = product_model(fx)
    fresh.load_state_dict(new_sd)
    c, _ = fresh.sample(context=dict(ctx), num_samples=B, num_sampling_steps=K, initial_noise=x_T, seed=5)
    assert not torch.equal(a, b) and torch.equal(b, c)


def test_num_sampling_steps_is_validated(golden):
    m = product_model(golden("c2"))
    with pytest.raises(ValueError):
        m.sample(context={"classes": torch.zeros(2, dtype=torch.long, device=DEV)}, num_samples=2,
                 num_sampling_steps=1001)


def test_seeded_shards_reproduce_the_unsharded_batch(golden):
    """dist.sample_sharded's contract without a process group: rows [lo, hi) sampled with row_offset = lo and the same
    seed equal those rows of the one-call batch (in-kernel Philox noise AND the seeded x_T), and differ from each other."""
    from xdiffusion_b200 import ops
    fx = golden("c2")
    m = product_model(fx)
    B, K = 16, 6
    cls = torch.arange(B, device=DEV) % 10
    ops.set_split_k(False)
    try:
        full, _ = m.sample(context={"classes": cls}, num_samples=B, num_sampling_steps=K, seed=21)
        parts = []
        for lo, hi in ((0, 8), (8, 16)):
            p, _ = m.sample(context={"classes": cls[lo:hi].contiguous()}, num_samples=hi - lo, num_sampling_steps=K,
                            seed=21, row_offset=lo)
            parts.append(p)
    finally:
        ops.set_split_k(True)
    assert torch.equal(torch.cat(parts), full) and not torch.equal(parts[0], parts[1])
